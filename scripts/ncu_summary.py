"""Turns one `ncu --set full --import-source on` report of a kernel into the two committed artefacts bench.py and DESIGN.md cite:
profiles/<tag>_traffic.json (DRAM bytes, FP64-pipe warp instructions, pipe / hit / stall figures, opcode mix) and a short text summary.
Runs where ncu is installed (no GPU needed):  python scripts/ncu_summary.py gpurun_out/r2b_sa.ncu-rep sparse_align_kernel 4096 r2b "<how the capture was taken>" """
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def export(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, kernel, pairs, tag, how = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
    raw = export(rep, "raw")
    m = {h: (u, v) for h, u, v in zip(raw[0], raw[1], raw[2])}

    def f(name, scale=None):
        u, v = m[name]
        v = float(v.replace(",", ""))
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        return v * mult.get(u, 1.0)

    src = export(rep, "source")
    hdr = src[1]; ix = {h: i for i, h in enumerate(hdr)}
    ops = {}
    for r in src[2:]:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]] or 0)
    total = sum(ops.values())
    fp64 = sum(ops.get(k, 0) for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    stall = lambda k: round(f("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k), 2)
    rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
    rec = {
        "pairs": pairs, "dram_bytes_read": rd, "dram_bytes_write": wr, "bytes_per_pair": (rd + wr) / pairs,
        "fp64_warp_insts": fp64, "fp64_warp_insts_per_pair": fp64 / pairs, "warp_instructions": total,
        "duration_ms_under_ncu": round(f("gpu__time_duration.sum"), 4),
        "issue_slots_busy_pct": round(100 * f("smsp__issue_active.avg.per_cycle_active"), 2),
        "fp64_pipe_pct": round(f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 2),
        "l1tex_throughput_pct": round(f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), 2),
        "l1tex_hit_pct": round(f("l1tex__t_sector_hit_rate.pct"), 2), "l2_hit_pct": round(f("lts__t_sector_hit_rate.pct"), 2),
        "dram_throughput_pct": round(f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 2),
        "registers": int(f("launch__registers_per_thread")),
        "stalls_per_issued_instruction": {k: stall(k) for k in ("long_scoreboard", "wait", "barrier", "short_scoreboard", "math_pipe_throttle", "not_selected", "dispatch_stall", "no_instruction")},
        "opcode_mix_pct": {k: round(100.0 * v / total, 1) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:12]},
        "source": how + "; FP64-pipe instruction count = executed DFMA + DMUL + DADD + DSETP warp instructions of the report's source page",
    }
    p = os.path.join(ROOT, "profiles", tag + "_traffic.json")
    json.dump({kernel: rec}, open(p, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
