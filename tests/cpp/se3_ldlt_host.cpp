// Host build of the product's register-only LDL^T / SE3 tail (dsdtm_b200/csrc/se3_ldlt.cuh) so the CPU suite can check
// it against the oracle and numpy without a GPU. Test infrastructure.
#include "../../dsdtm_b200/csrc/se3_ldlt.cuh"

extern "C" {
void prod_ldlt6_solve(const double* H /*36 row-major*/, const double* b, double* x)
{
    double Hm[6][6], bv[6], xv[6];
    for (int i = 0; i < 6; ++i) { bv[i] = b[i]; for (int j = 0; j < 6; ++j) Hm[i][j] = H[6 * i + j]; }
    dsdtm::ldlt6_solve_reg(Hm, bv, xv);
    for (int i = 0; i < 6; ++i) x[i] = xv[i];
}
// the form the sparse-alignment kernel's tail uses (DSDTM_SA_TAIL_CONST): the four series from the coefficient table, handed to se3_mul_exp
void prod_se3_mul_exp_table(const double* T, const double* x, double* out)
{
    static const double coef[32] = { DSDTM_SE3_SERIES_COEF_LIST };
    double Tv[7], xv[6], ov[7];
    for (int i = 0; i < 7; ++i) Tv[i] = T[i];
    for (int i = 0; i < 6; ++i) xv[i] = x[i];
    const double t2 = dsdtm::se3_theta2(xv);
    double s4[4] = { 0.0, 0.0, 0.0, 0.0 };
    if (dsdtm::se3_exp_uses_series(t2)) {
        const double h2 = 0.25 * t2;
        for (int j = 0; j < 4; ++j) {
            const double arg = (j < 2) ? h2 : t2;
            double r = coef[j];
            for (int k = 1; k < 8; ++k) r = fma(arg, r, coef[4 * k + j]);
            s4[j] = r;
        }
    }
    dsdtm::se3_mul_exp(Tv, xv, ov, s4);
    for (int i = 0; i < 7; ++i) out[i] = ov[i];
}
void prod_se3_mul_exp(const double* T, const double* x, double* out)
{
    double Tv[7], xv[6], ov[7];
    for (int i = 0; i < 7; ++i) Tv[i] = T[i];
    for (int i = 0; i < 6; ++i) xv[i] = x[i];
    dsdtm::se3_mul_exp(Tv, xv, ov);
    for (int i = 0; i < 7; ++i) out[i] = ov[i];
}
}
