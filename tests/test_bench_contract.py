"""CPU suite: the JSON contract of bench.py's reference arm (`--impl reference`: the oracle port of the path on the host cores,
rank 0 only) -- the keys the driver reads, for both camera geometries. The GPU arm's line is checked on the GPU box by the driver."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ); e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "8", "--scenes", "2"] + extra, capture_output=True, text=True, env=e, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


@pytest.mark.parametrize("cam,size", [("kinect", "640x480"), ("euroc", "752x480")])
def test_reference_arm_line(built, cam, size):
    lines = _run(["--cam", cam])
    assert len(lines) == 1                                              # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frame_pairs/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert size in d["metric"] and size in d["config"]["workload"] and d["config"]["camera"] == cam
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f64"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "8 pairs" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly(built):
    # under torchrun only rank 0 runs the CPU arm; the other ranks print nothing and exit 0
    lines = _run([], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert lines == []


def test_roofline_inputs_of_the_gpu_arm_are_committed():
    """bench.py's `roofline` for the dominant kernel reads the newest committed ncu capture (profiles/*_traffic.json): the FP64-pipe
    instruction count and the DRAM traffic per pair must be there and plausible, or the bench line silently loses its roofline."""
    import bench
    p = bench.traffic_file()
    assert os.path.exists(p), p
    fpi = bench.fp64_insts_per_pair("sparse_align_kernel")
    tr = bench.measured_traffic("sparse_align_kernel", 4096)
    assert fpi is not None and 5e4 < fpi < 2e5                      # ~78 k FP64 warp instructions per pair
    assert tr is not None and 1e9 < tr < 5e9                        # ~2.3 GB per launch of 4096 pairs
    s = bench.ncu_summary("sparse_align_kernel")
    assert "fp64_pipe_pct" in s and "source" in s
