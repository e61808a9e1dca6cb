// Test harness for cubature_b200/csrc/exact_acc.h (compiled by tests/test_exact_acc.py with g++): exposes the big-integer
// arithmetic of the exact running sums through a C ABI so that Python integers can check it.
#include "../cubature_b200/csrc/exact_acc.h"
#include "../cubature_b200/csrc/acc_window.h"
#include <vector>

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
// The rule kernels' path (acc_device.cuh): value i, with sign sgn[i], is added by "thread" i % T to its 256-bit window at
// `base` (or to the far cells when it does not fit); the windows are summed as 32-bit cells and folded into the totals the
// way k_exact_decide does.  Returns the signed sum as a double; *n_far = values that took the far path.
double xa_window_sum(const double* x, const int* sgn, int n, int T, int base, int nearest, int* n_far) {
    std::vector<xaw_u64> w((size_t)T * XA_WLIMBS, 0ULL);
    long long far[XA_LIMBS] = {0};
    int nf = 0;
    for (int i = 0; i < n; ++i) {
        const xa_u64 b = xa_bits(x[i]);
        bool neg = sgn[i] < 0;
        if (b >> 63) neg = !neg;
        const xa_u64 m = xa_mantissa(b);
        if (!m) continue;
        const int pos = xa_lsb_pos((int)((b >> 52) & 2047ULL));
        if (!xaw_accumulate(&w[(size_t)(i % T)], T, m, pos, base, neg)) {  // limb-major, like the shared-memory layout
            int j;
            long long c[3];
            xaw_far_cells(m, pos, neg, &j, c);
            far[j] += c[0];
            far[j + 1] += c[1];
            far[j + 2] += c[2];
            ++nf;
        }
    }
    if (n_far) *n_far = nf;
    long long cell[XA_HALVES] = {0};
    for (int t = 0; t < T; ++t)
        for (int h = 0; h < XA_HALVES; ++h) cell[h] += xaw_cell(w[(size_t)(h >> 1) * T + t], h);
    XaBig P, Q;
    P.clear();
    Q.clear();
    xa_fold_cells(&P, &Q, cell, base);
    for (int j = 0; j < XA_LIMBS - 3; ++j) xa_add_signed(&P, &Q, far[j], 32 * j);
    P.normalize();
    Q.normalize();
    if (P.cmp(Q) >= 0) {
        P.sub(Q);
        return P.to_double(nearest);
    }
    Q.sub(P);
    return -Q.to_double(nearest);
}
}
