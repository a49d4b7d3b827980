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
void prod_se3_mul_exp(const double* T, const double* x, double* out)
{
    double Tv[7], xv[6], ov[7];
    for (int i = 0; i < 7; ++i) Tv[i] = T[i];
    for (int i = 0; i < 6; ++i) xv[i] = x[i];
    dsdtm::se3_mul_exp(Tv, xv, ov);
    for (int i = 0; i < 7; ++i) out[i] = ov[i];
}
}
