// Test harness for cubature_b200/csrc/exact_acc.h (compiled by tests/test_exact_acc.py with g++): exposes the big-integer
// arithmetic of the exact running sums through a C ABI so that Python integers can check it.
#include "../cubature_b200/csrc/exact_acc.h"

extern "C" {
// sum of n non-negative finite doubles -> double, truncated (nearest = 0) or rounded to nearest even (1)
double xa_sum(const double* x, int n, int nearest) {
    XaBig b;
    b.clear();
    for (int i = 0; i < n; ++i) b.add_double(x[i]);
    b.normalize();
    return b.to_double(nearest);
}
// (sum of a) - (sum of b), requires sum a >= sum b
double xa_diff(const double* a, int na, const double* b, int nb, int nearest) {
    XaBig A, B;
    A.clear();
    B.clear();
    for (int i = 0; i < na; ++i) A.add_double(a[i]);
    for (int i = 0; i < nb; ++i) B.add_double(b[i]);
    A.normalize();
    B.normalize();
    if (A.cmp(B) < 0) return -1.0;
    A.sub(B);
    return A.to_double(nearest);
}
// the select kernel's residual: X = sum(x) with all popped values m_i * 2^pos taken out through the (H, L1, Ldbl) split
double xa_residual(const double* x, int n, int binade, const unsigned long long* popped_m, int np) {
    XaBig A, B;
    A.clear();
    for (int i = 0; i < n; ++i) A.add_double(x[i]);
    A.normalize();
    const int pos = xa_lsb_pos(binade);
    const xa_u128 H = ((xa_u128)A.window64(pos + 64) << 64) | A.window64(pos);
    const xa_u64 L1 = pos >= 64 ? A.window64(pos - 64) : (pos > 0 ? (A.window64(0) << (64 - pos)) : 0ULL);
    for (int j = 0; j < XA_LIMBS; ++j) {
        const int lo_bit = 32 * j;
        if (lo_bit >= pos)
            B.l[j] = 0ULL;
        else if (lo_bit + 32 <= pos)
            B.l[j] = A.l[j];
        else
            B.l[j] = A.l[j] & ((1ULL << (pos - lo_bit)) - 1ULL);
    }
    B.jlo = 0;
    B.jhi = XA_LIMBS - 1;
    const double Ldbl = B.to_double(0);
    xa_u128 P = 0;
    for (int i = 0; i < np; ++i) P += popped_m[i];
    const xa_u128 D = H - P;
    if (D == 0) return Ldbl;
    return xa_trunc_hi(D, L1, pos);
}
}
