// cubature_oracle.cpp -- ORACLE O1: a literal CPU restatement of the reference's h-adaptive loop.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under cubature_b200/ links, imports or executes this
// file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// The reference (jonathanschilling/Cubature, pure Java) cannot run here or on the GPU boxes: there
// is no JVM in the image.  This file restates, operation by operation, the
// reference classes it cites (paths relative to /root/reference/src/main/java/de/labathome/cubature):
//     Cubature.java:108-170, Rule.java:38-289, Rule75GenzMalik.java:28-326,
//     RuleGaussKronrod_1d.java:9-164, Region.java:17-55, RegionHeap.java:12-36, Hypercube.java:10-41
// plus the JDK's java.util.PriorityQueue (NOT under /root/reference; Java 11 target, semantics
// restated in SURVEY.md Appendix B: siftUp stops on compare >= 0, siftDown prefers the left child
// unless the right one is strictly "smaller", and stops on compare <= 0).
//
// Parity pin: the reference holds exactly one numeric known-answer test
// (TestCubature.java:435-465 == README.md:224-229: 13.69609043 +/- 0.00136919, each +-1e-8);
// tests/test_oracle.py checks this file against it (with glibc exp, as close to Math.exp as we can
// get, and with the deterministic exp of detmath.h).  Region counts, split dimensions, numEval,
// fdim > 1, the norms other than INDIVIDUAL, GK15 numerics and the variable transforms are NOT
// pinned by any reference test ("parity unpinned" for those; they are pinned only by this file).
//
// Deliberate deviations from the Java source (all marked "D" there):
//   * numEval / maxEval are int64 (Java int overflows at 2^31).
//   * the integrand is called on exactly nR*P points, not on the stale capacity-sized array.
//   * no stdout side effects (Rule.java:144-146, Rule.java:264).
//   * Math.pow(x, 1.5) in the GK15 error heuristic (RuleGaussKronrod_1d.java:147) is x*sqrt(x):
//     java.lang.Math.pow is only specified to 1 ulp, so no bit-exact restatement exists.
// Compile with -O2 -ffp-contract=off (Java never fuses multiply-add).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../cubature_b200/csrc/integrands.h"  // test INPUTS (integrand definitions), shared with the device

namespace orc {

// The reference's plugin boundary: double[fdim][n] apply(double[dim][n])  (Cubature.java:98-99).
// x[d] points to n coordinates of dimension d; out[j] receives n values of component j.
typedef void (*integrand_fn)(void* ctx, int dim, int64_t n, const double* const* x, int fdim, double* const* out);

struct ErrorEstimate {  // ErrorEstimate.java:3-6
    double val = 0.0, err = 0.0;
};

struct Hypercube {  // Hypercube.java:3-41
    int dim = 0;
    std::vector<double> centers, halfwidths;
    double volume = 0.0;

    void computeVolume() {  // Hypercube.java:36-41
        volume = 1.0;
        for (int i = 0; i < dim; ++i) volume *= halfwidths[i] * 2.0;
    }
    static Hypercube initFromRanges(const double* xmin, const double* xmax, int dim) {  // Hypercube.java:19-30
        Hypercube h;
        h.dim = dim;
        h.centers.resize(dim);
        h.halfwidths.resize(dim);
        for (int i = 0; i < dim; ++i) {
            h.centers[i] = (xmax[i] + xmin[i]) / 2.0;
            h.halfwidths[i] = (xmax[i] - xmin[i]) / 2.0;
        }
        h.computeVolume();
        return h;
    }
    Hypercube clone() const {  // Hypercube.java:10-17,32-34
        Hypercube h;
        h.dim = dim;
        h.centers = centers;
        h.halfwidths = halfwidths;
        h.computeVolume();
        return h;
    }
};

struct Region {  // Region.java:3-56
    Hypercube h;
    int splitDim = 0;
    int fdim = 0;
    std::vector<ErrorEstimate> ee;
    double errmax = 0.0;

    static Region* init(const Hypercube& h, int fdim) {  // Region.java:17-28
        Region* r = new Region();
        r->h = h.clone();
        r->splitDim = 0;
        r->fdim = fdim;
        r->ee.assign(fdim, ErrorEstimate());
        r->errmax = std::numeric_limits<double>::infinity();
        return r;
    }
    Region* cut() {  // Region.java:31-38
        h.halfwidths[splitDim] *= 0.5;
        h.volume *= 0.5;
        Region* r2 = Region::init(h.clone(), fdim);
        h.centers[splitDim] -= h.halfwidths[splitDim];
        r2->h.centers[splitDim] += h.halfwidths[splitDim];
        return r2;
    }
    int compareTo(const Region* o) const {  // Region.java:41-55
        if (errmax == o->errmax) return 0;
        return (errmax < o->errmax) ? 1 : -1;
    }
};

// Variant switch (orc_set_exact_sums; NOT reference behaviour, default off): the heap's running sums and the residual of
// the batch-pop loop are carried in binary128 (rounding-free for every problem size used here) and handed to
// Rule.converged as doubles -- values rounded to nearest, errors truncated toward zero, so that `err < tol` on the
// double is the exact comparison.  This is "the decision the reference would take if its incremental += / -= sums
// (RegionHeap.java:21-36, Rule.java:122-124) did not drift"; the device-side selection of libcubature_cuda computes
// exactly that (integer accumulators), and comparing the two variants shows where the reference's own rounding decides.
static int g_exact_sums = 0;
static double q_to_double_trunc(__float128 x) {
    double d = (double)x;
    if ((__float128)d > x && x >= 0) d = std::nextafter(d, 0.0);
    if ((__float128)d < x && x < 0) d = std::nextafter(d, 0.0);
    return d;
}

// RegionHeap.java:5-37 on top of java.util.PriorityQueue.
struct RegionHeap {
    int fdim;
    std::vector<ErrorEstimate> ee;
    std::vector<__float128> qval, qerr;  // exact-sums variant only
    std::vector<Region*> q;  // the PriorityQueue's backing array, q[0..size)

    explicit RegionHeap(int fdim_) : fdim(fdim_), ee(fdim_), qval(fdim_, 0), qerr(fdim_, 0) {}
    void sync_exact() {  // exact-sums variant: the double sums are roundings of the binary128 ones
        for (int i = 0; i < fdim; ++i) {
            ee[i].val = (double)qval[i];
            ee[i].err = q_to_double_trunc(qerr[i]);
        }
    }
    size_t size() const { return q.size(); }

    void add(Region* e) {  // RegionHeap.java:30-36 then PriorityQueue.offer/siftUp
        for (int i = 0; i < fdim; ++i) {
            ee[i].val += e->ee[i].val;
            ee[i].err += e->ee[i].err;
        }
        if (g_exact_sums) {
            for (int i = 0; i < fdim; ++i) {
                qval[i] += e->ee[i].val;
                qerr[i] += e->ee[i].err;
            }
            sync_exact();
        }
        size_t k = q.size();
        q.push_back(nullptr);
        while (k > 0) {
            size_t parent = (k - 1) >> 1;
            Region* p = q[parent];
            if (e->compareTo(p) >= 0) break;
            q[k] = p;
            k = parent;
        }
        q[k] = e;
    }
    Region* poll() {  // PriorityQueue.poll/siftDown then RegionHeap.java:21-28
        Region* result = q[0];
        size_t n = q.size() - 1;
        Region* x = q[n];
        q.pop_back();
        if (n > 0) {
            size_t k = 0, half = n >> 1;
            while (k < half) {
                size_t child = (k << 1) + 1;
                Region* c = q[child];
                size_t right = child + 1;
                if (right < n && c->compareTo(q[right]) > 0) c = q[child = right];
                if (x->compareTo(c) <= 0) break;
                q[k] = c;
                k = child;
            }
            q[k] = x;
        }
        for (int i = 0; i < fdim; ++i) {
            ee[i].val -= result->ee[i].val;
            ee[i].err -= result->ee[i].err;
        }
        if (g_exact_sums) {
            for (int i = 0; i < fdim; ++i) {
                qval[i] -= result->ee[i].val;
                qerr[i] -= result->ee[i].err;
            }
            sync_exact();
        }
        return result;
    }
};

enum { NORM_INDIVIDUAL = 0, NORM_PAIRED = 1, NORM_L2 = 2, NORM_L1 = 3, NORM_LINF = 4 };  // CubatureError.java:9-29

static double errMax(int fdim, const std::vector<ErrorEstimate>& ee) {  // Rule.java:281-289
    double errmax = 0.0;
    for (int k = 0; k < fdim; ++k)
        if (ee[k].err > errmax) errmax = ee[k].err;
    return errmax;
}

// Rule.java:162-279.  Returns -1 for "both tolerances NaN" (the Java code throws).
// EXTENSION, not reference behaviour: norm | 8 on INDIVIDUAL / PAIRED requires EVERY
// component (pair) to meet a tolerance, as README.md:112-122 documents and upstream C cubature implements.
static int converged(int fdim, const std::vector<ErrorEstimate>& ee, double absTol, double relTol, int norm) {
    if (std::isnan(relTol) && std::isnan(absTol)) return -1;
    bool conv = false;
    int j;
    if ((norm & 8) && (norm & 7) <= 1) {
        conv = true;
        if ((norm & 7) == 0) {
            for (j = 0; j < fdim; ++j) {
                bool cj = false;
                if (!std::isnan(absTol)) cj |= ee[j].err < absTol;
                if (!std::isnan(relTol)) cj |= ee[j].err < std::fabs(ee[j].val) * relTol;
                conv &= cj;
            }
        } else {
            for (j = 0; j + 1 < fdim; j += 2) {
                double maxerr = ee[j].err > ee[j + 1].err ? ee[j].err : ee[j + 1].err;
                double maxval = ee[j].val > ee[j + 1].val ? ee[j].val : ee[j + 1].val;
                double serr = maxerr > 0 ? 1 / maxerr : 1;
                double sval = maxval > 0 ? 1 / maxval : 1;
                double err = std::sqrt((ee[j].err * serr) * (ee[j].err * serr) + (ee[j + 1].err * serr) * (ee[j + 1].err * serr)) * maxerr;
                double val = std::sqrt((ee[j].val * sval) * (ee[j].val * sval) + (ee[j + 1].val * sval) * (ee[j + 1].val * sval)) * maxval;
                bool cj = false;
                if (!std::isnan(absTol)) cj |= err < absTol;
                if (!std::isnan(relTol)) cj |= err < std::fabs(val) * relTol;
                conv &= cj;
            }
            if (j < fdim) {
                bool cj = false;
                if (!std::isnan(absTol)) cj |= ee[j].err < absTol;
                if (!std::isnan(relTol)) cj |= ee[j].err < std::fabs(ee[j].val) * relTol;
                conv &= cj;
            }
        }
        return conv ? 1 : 0;
    }
    switch (norm & 7) {
        case NORM_INDIVIDUAL:
            for (j = 0; j < fdim; ++j) {
                if (!std::isnan(absTol)) conv |= ee[j].err < absTol;
                if (!std::isnan(relTol)) conv |= ee[j].err < std::fabs(ee[j].val) * relTol;
            }
            break;
        case NORM_PAIRED:
            for (j = 0; j + 1 < fdim; j += 2) {
                double maxerr, serr, err, maxval, sval, val;
                maxerr = ee[j].err > ee[j + 1].err ? ee[j].err : ee[j + 1].err;
                maxval = ee[j].val > ee[j + 1].val ? ee[j].val : ee[j + 1].val;
                serr = maxerr > 0 ? 1 / maxerr : 1;
                sval = maxval > 0 ? 1 / maxval : 1;
                err = std::sqrt((ee[j].err * serr) * (ee[j].err * serr) + (ee[j + 1].err * serr) * (ee[j + 1].err * serr)) * maxerr;
                val = std::sqrt((ee[j].val * sval) * (ee[j].val * sval) + (ee[j + 1].val * sval) * (ee[j + 1].val * sval)) * maxval;
                if (!std::isnan(absTol)) conv |= err < absTol;
                if (!std::isnan(relTol)) conv |= err < std::fabs(val) * relTol;
            }
            if (j < fdim) {
                if (!std::isnan(absTol)) conv |= ee[j].err < absTol;
                if (!std::isnan(relTol)) conv |= ee[j].err < std::fabs(ee[j].val) * relTol;
            }
            break;
        case NORM_L1: {
            double err = 0, val = 0;
            for (j = 0; j < fdim; ++j) {
                err += ee[j].err;
                val += std::fabs(ee[j].val);
            }
            if (!std::isnan(absTol)) conv |= err < absTol;
            if (!std::isnan(relTol)) conv |= err < val * relTol;
        } break;
        case NORM_LINF: {
            double err = 0, val = 0;
            for (j = 0; j < fdim; ++j) {
                double absval = std::fabs(ee[j].val);
                if (ee[j].err > err) err = ee[j].err;
                if (absval > val) val = absval;
            }
            if (!std::isnan(absTol)) conv |= err < absTol;
            if (!std::isnan(relTol)) conv |= err < val * relTol;
        } break;
        case NORM_L2: {
            double maxerr = 0, maxval = 0, serr, sval, err = 0, val = 0;
            for (j = 0; j < fdim; ++j) {
                double absval = std::fabs(ee[j].val);
                if (ee[j].err > maxerr) maxerr = ee[j].err;
                if (absval > maxval) maxval = absval;
            }
            serr = maxerr > 0 ? 1 / maxerr : 1;
            sval = maxval > 0 ? 1 / maxval : 1;
            for (j = 0; j < fdim; ++j) {
                err += (ee[j].err * serr) * (ee[j].err * serr);
                val += (std::fabs(ee[j].val) * sval) * (std::fabs(ee[j].val) * sval);
            }
            err = std::sqrt(err) * maxerr;
            val = std::sqrt(val) * maxval;
            if (!std::isnan(absTol)) conv |= err < absTol;
            if (!std::isnan(relTol)) conv |= err < val * relTol;
        } break;
        default:
            conv = false;
    }
    return conv ? 1 : 0;
}

struct Trace {
    std::vector<int64_t> batch;  // nR of every evalError call after the first (Rule.java:135)
};

static int g_threads = 1;  // threads evaluating one integrand batch (orc_set_threads)

struct Rule {  // Rule.java:6-49
    int dim, fdim, num_points;
    std::vector<std::vector<double>> pts;   // [dim][nR*num_points]
    std::vector<std::vector<double>> vals;  // [fdim][nR*num_points]
    integrand_fn fn;
    void* ctx;

    Rule(int dim_, int fdim_, int np) : dim(dim_), fdim(fdim_), num_points(np), fn(nullptr), ctx(nullptr) {}
    virtual ~Rule() {}
    virtual void evalError(Region** R, int64_t nR) = 0;

    void alloc_rule_pts(int64_t nR) {  // Rule.java:38-49 without the 2x over-allocation (App. C: D)
        pts.assign(dim, std::vector<double>((size_t)(nR * num_points)));
        vals.assign(fdim, std::vector<double>((size_t)(nR * num_points)));
    }
    // ONE integrand call for all points of the batch (Rule75GenzMalik.java:108, RuleGaussKronrod_1d.java:88).  The
    // reference is single-threaded; its README (129-163) points out that this batch interface lets the USER evaluate
    // the points in parallel.  orc_set_threads(T > 1) does exactly that -- the batch is cut into T slices evaluated by T
    // threads (pointwise integrand: identical bits) -- and nothing else: point generation, rule sums and the heap stay
    // serial as in the reference.  Used only by bench.py's CPU baseline; the parity tests run with T = 1.
    void call_integrand(int64_t n) {
        const int T = g_threads;
        if (T <= 1 || n < 4096) {
            std::vector<const double*> xp(dim);
            std::vector<double*> fp(fdim);
            for (int d = 0; d < dim; ++d) xp[d] = pts[d].data();
            for (int j = 0; j < fdim; ++j) fp[j] = vals[j].data();
            fn(ctx, dim, n, xp.data(), fdim, fp.data());
            return;
        }
        std::vector<std::thread> pool;
        const int64_t per = (n + T - 1) / T;
        for (int t = 0; t < T; ++t) {
            const int64_t lo = per * t, hi = std::min<int64_t>(n, lo + per);
            if (lo >= hi) break;
            pool.emplace_back([this, lo, hi] {
                std::vector<const double*> xp(dim);
                std::vector<double*> fp(fdim);
                for (int d = 0; d < dim; ++d) xp[d] = pts[d].data() + lo;
                for (int j = 0; j < fdim; ++j) fp[j] = vals[j].data() + lo;
                fn(ctx, dim, hi - lo, xp.data(), fdim, fp.data());
            });
        }
        for (std::thread& th : pool) th.join();
    }

    // Rule.java:51-160.  Returns 0, or -1 if both tolerances are NaN.
    int cubature(int64_t maxEval, double relTol, double absTol, double* val, double* err, const Hypercube& h,
                 int norm, RegionHeap& regions, int64_t& numEvalOut, int& convergedOut, Trace& trace) {
        int64_t numEval = 0;
        std::vector<Region*> R(2, nullptr);
        R[0] = Region::init(h, fdim);
        {
            Region* one[1] = {R[0]};
            evalError(one, 1);
        }
        R[0]->errmax = errMax(R[0]->fdim, R[0]->ee);
        regions.add(R[0]);
        numEval += num_points;

        bool conv = false;
        while (numEval < maxEval || maxEval == 0) {
            int c = converged(fdim, regions.ee, absTol, relTol, norm);
            if (c < 0) return -1;
            if (c) {
                conv = true;
                break;
            }
            int64_t nR = 0;
            std::vector<ErrorEstimate> ee(fdim);
            for (int j = 0; j < fdim; ++j) {
                ee[j].err = regions.ee[j].err;
                ee[j].val = regions.ee[j].val;
            }
            std::vector<__float128> qres;
            if (g_exact_sums) qres = regions.qerr;
            do {
                if ((size_t)(nR + 2) > R.size()) R.resize((size_t)(nR + 2) * 2, nullptr);
                R[nR] = regions.poll();
                for (int j = 0; j < fdim; ++j) ee[j].err -= R[nR]->ee[j].err;
                if (g_exact_sums)
                    for (int j = 0; j < fdim; ++j) {
                        qres[j] -= R[nR]->ee[j].err;
                        ee[j].err = q_to_double_trunc(qres[j]);
                    }
                R[nR + 1] = R[nR]->cut();
                numEval += (int64_t)num_points * 2;
                nR += 2;
                if (converged(fdim, ee, absTol, relTol, norm) > 0) break;
            } while (regions.size() > 0 && (numEval < maxEval || maxEval == 0));

            evalError(R.data(), nR);
            trace.batch.push_back(nR);
            for (int64_t i = 0; i < nR; ++i) {
                R[i]->errmax = errMax(R[i]->fdim, R[i]->ee);
                regions.add(R[i]);
            }
        }
        for (int j = 0; j < fdim; ++j) val[j] = err[j] = 0;
        for (size_t i = 0; i < regions.q.size(); ++i) {  // heap ARRAY order (Rule.java:153-159)
            for (int j = 0; j < fdim; ++j) {
                val[j] += regions.q[i]->ee[j].val;
                err[j] += regions.q[i]->ee[j].err;
            }
        }
        numEvalOut = numEval;
        convergedOut = conv ? 1 : 0;
        return 0;
    }
};

// Rule75GenzMalik.java:15-326
struct Rule75GenzMalik : Rule {
    std::vector<double> widthLambda, widthLambda2, p;
    double weight2, weight4, weightE2, weightE4;
    double weight1, weight3, weight5, weightE1, weightE3;

    static int num0_0(int) { return 1; }
    static int numR0_0fs(int dim) { return 2 * dim; }
    static int numRR0_0fs(int dim) { return 2 * dim * (dim - 1); }
    static int numR_Rfs(int dim) { return 1 << dim; }
    static int numPoints(int dim) {  // Cubature.java:158-162
        return num0_0(dim) + 2 * numR0_0fs(dim) + numRR0_0fs(dim) + numR_Rfs(dim);
    }

    Rule75GenzMalik(int dim_, int fdim_) : Rule(dim_, fdim_, numPoints(dim_)) {  // Rule75GenzMalik.java:28-58
        int dim = dim_;
        weight1 = (12824 - 9120 * dim + 400 * dim * dim) / 19683.0;  // int arithmetic, then one division
        weight3 = (1820 - 400 * dim) / 19683.0;
        weight5 = 6859.0 / (19683.0 * (1 << dim));
        weightE1 = (729 - 950 * dim + 50 * dim * dim) / 729.0;
        weightE3 = (265 - 100 * dim) / 1458.0;
        weight2 = 980.0 / 6561.0;
        weight4 = 200.0 / 19683.0;
        weightE2 = 245.0 / 486.0;
        weightE4 = 25.0 / 729.0;
        p.resize(dim);
        widthLambda.resize(dim);
        widthLambda2.resize(dim);
    }

    static int ls0(int n) {  // Rule75GenzMalik.java:196-223: index of the least-significant 0 bit
        int bit = 0;
        while (n & 1) {
            n >>= 1;
            ++bit;
        }
        return bit;
    }
    void store(int64_t off) {
        for (int d = 0; d < dim; ++d) pts[d][(size_t)off] = p[d];
    }
    void evalR_Rfs(int64_t off, const double* c, const double* r) {  // Rule75GenzMalik.java:237-269
        int signs = 0;
        for (int i = 0; i < dim; ++i) p[i] = c[i] + r[i];
        for (int i = 0;; ++i) {
            store(off++);
            int d = ls0(i);
            if (d >= dim) break;
            int mask = 1 << d;
            signs ^= mask;
            p[d] = (signs & mask) != 0 ? c[d] - r[d] : c[d] + r[d];
        }
    }
    void evalRR0_0fs(int64_t off, const double* c, const double* r) {  // Rule75GenzMalik.java:271-297
        for (int i = 0; i < dim - 1; ++i) {
            p[i] = c[i] - r[i];
            for (int j = i + 1; j < dim; ++j) {
                p[j] = c[j] - r[j];
                store(off++);
                p[i] = c[i] + r[i];
                store(off++);
                p[j] = c[j] + r[j];
                store(off++);
                p[i] = c[i] - r[i];
                store(off++);
                p[j] = c[j];
            }
            p[i] = c[i];
        }
    }
    void evalR0_0fs4d(int64_t off, const double* c, const double* r1, const double* r2) {  // Rule75GenzMalik.java:299-326
        store(off++);
        for (int i = 0; i < dim; ++i) {
            p[i] = c[i] - r1[i];
            store(off++);
            p[i] = c[i] + r1[i];
            store(off++);
            p[i] = c[i] - r2[i];
            store(off++);
            p[i] = c[i] + r2[i];
            store(off++);
            p[i] = c[i];
        }
    }

    void evalError(Region** R, int64_t nR) override {  // Rule75GenzMalik.java:61-184
        alloc_rule_pts(nR);
        const double lambda2 = 0.3585685828003180919906451539079374954541;
        const double lambda4 = 0.9486832980505137995996680633298155601160;
        const double lambda5 = 0.6882472016116852977216287342936235251269;
        const double ratio = (lambda2 * lambda2) / (lambda4 * lambda4);

        int64_t npts = 0;
        for (int64_t iR = 0; iR < nR; ++iR) {
            const double* center = R[iR]->h.centers.data();
            const double* halfwidth = R[iR]->h.halfwidths.data();
            for (int i = 0; i < dim; ++i) p[i] = center[i];
            for (int i = 0; i < dim; ++i) widthLambda2[i] = halfwidth[i] * lambda2;
            for (int i = 0; i < dim; ++i) widthLambda[i] = halfwidth[i] * lambda4;
            evalR0_0fs4d(npts, center, widthLambda2.data(), widthLambda.data());
            npts += num0_0(dim) + 2 * numR0_0fs(dim);
            evalRR0_0fs(npts, center, widthLambda.data());
            npts += numRR0_0fs(dim);
            for (int i = 0; i < dim; ++i) widthLambda[i] = halfwidth[i] * lambda5;
            evalR_Rfs(npts, center, widthLambda.data());
            npts += numR_Rfs(dim);
        }
        call_integrand(npts);

        std::vector<std::vector<double>> diff(dim, std::vector<double>((size_t)nR, 0.0));
        std::vector<double> v(num_points);
        for (int j = 0; j < fdim; ++j) {
            for (int64_t iR = 0; iR < nR; ++iR) {
                std::memcpy(v.data(), &vals[j][(size_t)(iR * num_points)], sizeof(double) * num_points);
                double result, res5th;
                double val0, sum2 = 0, sum3 = 0, sum4 = 0, sum5 = 0;
                int k, k0 = 0;
                val0 = v[0];
                k0 += 1;
                for (k = 0; k < dim; ++k) {
                    double v0 = v[(k0 + 4 * k) + 0];
                    double v1 = v[(k0 + 4 * k) + 1];
                    double v2 = v[(k0 + 4 * k) + 2];
                    double v3 = v[(k0 + 4 * k) + 3];
                    sum2 += v0 + v1;
                    sum3 += v2 + v3;
                    diff[k][(size_t)iR] += std::fabs(v0 + v1 - 2 * val0 - ratio * (v2 + v3 - 2 * val0));
                }
                k0 += 4 * k;
                for (k = 0; k < numRR0_0fs(dim); ++k) sum4 += v[k0 + k];
                k0 += k;
                for (k = 0; k < numR_Rfs(dim); ++k) sum5 += v[k0 + k];
                result = R[iR]->h.volume * (weight1 * val0 + weight2 * sum2 + weight3 * sum3 + weight4 * sum4 + weight5 * sum5);
                res5th = R[iR]->h.volume * (weightE1 * val0 + weightE2 * sum2 + weightE3 * sum3 + weightE4 * sum4);
                R[iR]->ee[j].val = result;
                R[iR]->ee[j].err = std::fabs(res5th - result);
            }
        }
        for (int64_t iR = 0; iR < nR; ++iR) {  // Rule75GenzMalik.java:172-183
            double maxdiff = 0;
            int dimDiffMax = 0;
            for (int i = 0; i < dim; ++i) {
                if (diff[i][(size_t)iR] > maxdiff) {
                    maxdiff = diff[i][(size_t)iR];
                    dimDiffMax = i;
                }
            }
            R[iR]->splitDim = dimDiffMax;
        }
    }
};

// RuleGaussKronrod_1d.java:9-164
struct RuleGaussKronrod_1d : Rule {
    static constexpr int n = 8;
    RuleGaussKronrod_1d(int dim_, int fdim_) : Rule(dim_, fdim_, 15) {}

    void evalError(Region** R, int64_t nR) override {
        static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                                      0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                                      0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                      0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
        static const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                                     0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
        static const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                                      0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                                      0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                      0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
        const double DBL_EPS = 2.220446049250313e-16;       // Math.ulp(1.0)
        const double MIN_NORMAL = 2.2250738585072014e-308;  // Double.MIN_NORMAL
        alloc_rule_pts(nR);
        for (int64_t iR = 0; iR < nR; ++iR) {
            double center = R[iR]->h.centers[0];
            double halfwidth = R[iR]->h.halfwidths[0];
            int npts = 0;
            pts[0][(size_t)(iR * 15 + npts++)] = center;
            for (int j = 1; j < n - 1; j += 2) {
                double w = halfwidth * xgk[j];
                pts[0][(size_t)(iR * 15 + npts++)] = center - w;
                pts[0][(size_t)(iR * 15 + npts++)] = center + w;
            }
            for (int j = 0; j < n; j += 2) {
                double w = halfwidth * xgk[j];
                pts[0][(size_t)(iR * 15 + npts++)] = center - w;
                pts[0][(size_t)(iR * 15 + npts++)] = center + w;
            }
            R[iR]->splitDim = 0;
        }
        call_integrand(nR * 15);
        for (int k = 0; k < fdim; ++k) {
            const double* vk = vals[k].data();
            for (int64_t iR = 0; iR < nR; ++iR) {
                const double* q = vk + iR * 15;
                double halfwidth = R[iR]->h.halfwidths[0];
                double result_gauss = q[0] * wg[n / 2 - 1];
                double result_kronrod = q[0] * wgk[n - 1];
                double result_abs = std::fabs(result_kronrod);
                double result_asc, mean, err;
                int npts = 1;
                for (int j = 0; j < (n - 1) / 2; ++j) {
                    int j2 = 2 * j + 1;
                    double v = q[npts] + q[npts + 1];
                    result_gauss += wg[j] * v;
                    result_kronrod += wgk[j2] * v;
                    result_abs += wgk[j2] * (std::fabs(q[npts]) + std::fabs(q[npts + 1]));
                    npts += 2;
                }
                for (int j = 0; j < n / 2; ++j) {
                    int j2 = 2 * j;
                    result_kronrod += wgk[j2] * (q[npts] + q[npts + 1]);
                    result_abs += wgk[j2] * (std::fabs(q[npts]) + std::fabs(q[npts + 1]));
                    npts += 2;
                }
                R[iR]->ee[k].val = result_kronrod * halfwidth;
                mean = result_kronrod * 0.5;
                result_asc = wgk[n - 1] * std::fabs(q[0] - mean);
                npts = 1;
                for (int j = 0; j < (n - 1) / 2; ++j) {
                    int j2 = 2 * j + 1;
                    result_asc += wgk[j2] * (std::fabs(q[npts] - mean) + std::fabs(q[npts + 1] - mean));
                    npts += 2;
                }
                for (int j = 0; j < n / 2; ++j) {
                    int j2 = 2 * j;
                    result_asc += wgk[j2] * (std::fabs(q[npts] - mean) + std::fabs(q[npts + 1] - mean));
                    npts += 2;
                }
                err = std::fabs(result_kronrod - result_gauss) * halfwidth;
                result_abs *= halfwidth;
                result_asc *= halfwidth;
                if (result_asc != 0 && err != 0) {
                    double base = 200 * err / result_asc;
                    double scale = base * std::sqrt(base);  // stands in for Math.pow(base, 1.5), see header
                    err = (scale < 1) ? result_asc * scale : result_asc;
                }
                if (result_abs > MIN_NORMAL / (50 * DBL_EPS)) {
                    double min_err = 50 * DBL_EPS * result_abs;
                    if (min_err > err) err = min_err;
                }
                R[iR]->ee[k].err = err;
            }
        }
        for (int64_t iR = 0; iR < nR; ++iR) R[iR]->errmax = errMax(fdim, R[iR]->ee);
    }
};

// ---- integrand plugins for the oracle -----------------------------------------------------------

struct BuiltinCtx {
    int family, dim, fdim;
    std::vector<double> params;
    int transform[32];
    double shift[32];
    bool any_transform;
    int use_libm;  // family 10 only: std::exp instead of dm_exp (KAT check against "a" libm)
};

template <class F, int DIM>
static void run_family(const BuiltinCtx* c, int64_t n, const double* const* x, double* const* out) {
    F functor(c->params.data(), c->fdim);
    std::vector<double> f((size_t)c->fdim);
    double pt[DIM];
    for (int64_t i = 0; i < n; ++i) {
        double jac = 1.0;
        for (int d = 0; d < DIM; ++d) {
            if (c->any_transform)
                jac = jac * cub_transform_coord(c->transform[d], c->shift[d], x[d][i], &pt[d]);
            else
                pt[d] = x[d][i];
        }
        functor.eval(pt, f.data());
        for (int j = 0; j < c->fdim; ++j) out[j][i] = c->any_transform ? f[(size_t)j] * jac : f[(size_t)j];
    }
}

template <int DIM>
static bool dispatch_family(const BuiltinCtx* c, int64_t n, const double* const* x, double* const* out) {
    switch (c->family) {
        case CUB_FAM_GENZ_OSCILLATORY: run_family<GenzOscillatory<DIM>, DIM>(c, n, x, out); return true;
        case CUB_FAM_GENZ_PRODUCT_PEAK: run_family<GenzProductPeak<DIM>, DIM>(c, n, x, out); return true;
        case CUB_FAM_GENZ_CORNER_PEAK: run_family<GenzCornerPeak<DIM>, DIM>(c, n, x, out); return true;
        case CUB_FAM_GENZ_GAUSSIAN: run_family<GenzGaussian<DIM>, DIM>(c, n, x, out); return true;
        case CUB_FAM_GENZ_C0: run_family<GenzC0<DIM>, DIM>(c, n, x, out); return true;
        case CUB_FAM_GENZ_DISCONTINUOUS: run_family<GenzDiscontinuous<DIM>, DIM>(c, n, x, out); return true;
        case CUB_FAM_GAUSS_ISO:
            if (c->use_libm) {
                for (int64_t i = 0; i < n; ++i) {
                    double s = 0.0;
                    for (int d = 0; d < DIM; ++d) s += x[d][i] * x[d][i];
                    out[0][i] = std::exp(-c->params[0] * s);
                }
            } else {
                run_family<GaussIso<DIM>, DIM>(c, n, x, out);
            }
            return true;
        case CUB_FAM_REFTEST: run_family<RefTest<DIM>, DIM>(c, n, x, out); return true;
        case CUB_FAM_OSC_VEC: run_family<OscVec<DIM>, DIM>(c, n, x, out); return true;
        case CUB_FAM_CEXP: run_family<CExp<DIM>, DIM>(c, n, x, out); return true;
    }
    return false;
}

static void builtin_integrand(void* vctx, int dim, int64_t n, const double* const* x, int, double* const* out) {
    const BuiltinCtx* c = (const BuiltinCtx*)vctx;
    bool ok = false;
    switch (dim) {
        case 1: ok = dispatch_family<1>(c, n, x, out); break;
        case 2: ok = dispatch_family<2>(c, n, x, out); break;
        case 3: ok = dispatch_family<3>(c, n, x, out); break;
        case 4: ok = dispatch_family<4>(c, n, x, out); break;
        case 5: ok = dispatch_family<5>(c, n, x, out); break;
        case 6: ok = dispatch_family<6>(c, n, x, out); break;
        case 7: ok = dispatch_family<7>(c, n, x, out); break;
        case 8: ok = dispatch_family<8>(c, n, x, out); break;
        case 9: ok = dispatch_family<9>(c, n, x, out); break;
        case 10: ok = dispatch_family<10>(c, n, x, out); break;
    }
    if (!ok) {
        std::fprintf(stderr, "oracle: unsupported family %d / dim %d\n", c->family, dim);
        std::abort();
    }
}

// A user plugin compiled by g++ from the SAME source string that NVRTC compiles for the device:
//   void entry(const double* x, double* f, const double* p)     x[dim], f[fdim]
typedef void (*user_point_fn)(const double* x, double* f, const double* p);
struct UserCtx {
    user_point_fn fn;
    int dim, fdim;
    std::vector<double> params;
    int transform[32];
    double shift[32];
    bool any_transform;
};
static void user_integrand(void* vctx, int dim, int64_t n, const double* const* x, int fdim, double* const* out) {
    const UserCtx* c = (const UserCtx*)vctx;
    std::vector<double> f((size_t)fdim), pt((size_t)dim);
    for (int64_t i = 0; i < n; ++i) {
        double jac = 1.0;
        for (int d = 0; d < dim; ++d) {
            if (c->any_transform)
                jac = jac * cub_transform_coord(c->transform[d], c->shift[d], x[d][i], &pt[(size_t)d]);
            else
                pt[(size_t)d] = x[d][i];
        }
        c->fn(pt.data(), f.data(), c->params.data());
        for (int j = 0; j < fdim; ++j) out[j][i] = c->any_transform ? f[(size_t)j] * jac : f[(size_t)j];
    }
}

struct Result {
    int status = 0;
    int dim = 0, fdim = 0;
    int64_t numEval = 0;
    int converged = 0;
    std::vector<double> val, err;
    Trace trace;
    // final regions in heap-array order
    std::vector<double> centers, halfwidths, rval, rerr, errmax;
    std::vector<int> splitDim;
};

// Cubature.java:108-170 (fdim is declared instead of probed at the box centre: App. C "D")
static Result* integrate(integrand_fn fn, void* ctx, int dim, int fdim, const double* xmin, const double* xmax,
                         double relTol, double absTol, int norm, int64_t maxEval) {
    Result* res = new Result();
    res->dim = dim;
    res->fdim = fdim;
    if (dim <= 0 || fdim <= 0) {
        res->status = -2;
        return res;
    }
    res->val.assign(fdim, 0.0);
    res->err.assign(fdim, std::numeric_limits<double>::infinity());
    Hypercube h = Hypercube::initFromRanges(xmin, xmax, dim);
    Rule* r;
    if (dim == 1) {
        r = new RuleGaussKronrod_1d(dim, fdim);
    } else {
        if (dim >= 32) {  // Rule75GenzMalik.java:35-42
            res->status = -3;
            return res;
        }
        r = new Rule75GenzMalik(dim, fdim);
    }
    r->fn = fn;
    r->ctx = ctx;
    RegionHeap regions(fdim);
    int st = r->cubature(maxEval, relTol, absTol, res->val.data(), res->err.data(), h, norm, regions, res->numEval,
                         res->converged, res->trace);
    res->status = st;
    size_t n = regions.q.size();
    res->centers.resize(n * dim);
    res->halfwidths.resize(n * dim);
    res->rval.resize(n * fdim);
    res->rerr.resize(n * fdim);
    res->errmax.resize(n);
    res->splitDim.resize(n);
    for (size_t i = 0; i < n; ++i) {
        Region* g = regions.q[i];
        for (int d = 0; d < dim; ++d) {
            res->centers[i * dim + d] = g->h.centers[d];
            res->halfwidths[i * dim + d] = g->h.halfwidths[d];
        }
        for (int j = 0; j < fdim; ++j) {
            res->rval[i * fdim + j] = g->ee[j].val;
            res->rerr[i * fdim + j] = g->ee[j].err;
        }
        res->errmax[i] = g->errmax;
        res->splitDim[i] = g->splitDim;
    }
    for (size_t i = 0; i < n; ++i) delete regions.q[i];
    delete r;
    return res;
}

static void fill_transform(int dim, const int* transform, const double* xmin, const double* xmax, int* tr, double* shift,
                           bool* any, double* tmin, double* tmax) {
    *any = false;
    for (int d = 0; d < dim; ++d) {
        int code = transform ? transform[d] : 0;
        tr[d] = code;
        shift[d] = 0.0;
        tmin[d] = xmin[d];
        tmax[d] = xmax[d];
        if (code == CUB_TR_SEMI_INF_UP) {
            shift[d] = xmin[d];
            tmin[d] = 0.0;
            tmax[d] = 1.0;
        } else if (code == CUB_TR_SEMI_INF_DOWN) {
            shift[d] = xmax[d];
            tmin[d] = 0.0;
            tmax[d] = 1.0;
        } else if (code == CUB_TR_INF) {
            tmin[d] = -1.0;
            tmax[d] = 1.0;
        }
        if (code != 0) *any = true;
    }
}

}  // namespace orc

// ---- C interface for tests / bench (ctypes) -------------------------------------------------------
extern "C" {

void orc_set_threads(int t) { orc::g_threads = t < 1 ? 1 : t; }
void orc_set_exact_sums(int on) { orc::g_exact_sums = on ? 1 : 0; }

void* orc_integrate_builtin(int family, int dim, int fdim, const double* params, int nparams, const double* xmin,
                            const double* xmax, const int* transform, double relTol, double absTol, int norm,
                            long long maxEval, int use_libm) {
    orc::BuiltinCtx c;
    c.family = family;
    c.dim = dim;
    c.fdim = fdim;
    c.params.assign(params, params + nparams);
    c.use_libm = use_libm;
    double tmin[32], tmax[32];
    if (dim > 32 || dim < 1) return nullptr;
    orc::fill_transform(dim, transform, xmin, xmax, c.transform, c.shift, &c.any_transform, tmin, tmax);
    return orc::integrate(orc::builtin_integrand, &c, dim, fdim, tmin, tmax, relTol, absTol, norm, maxEval);
}

void* orc_integrate_user(void* point_fn, int dim, int fdim, const double* params, int nparams, const double* xmin,
                         const double* xmax, const int* transform, double relTol, double absTol, int norm,
                         long long maxEval) {
    orc::UserCtx c;
    c.fn = (orc::user_point_fn)point_fn;
    c.dim = dim;
    c.fdim = fdim;
    c.params.assign(params, params + nparams);
    double tmin[32], tmax[32];
    if (dim > 32 || dim < 1) return nullptr;
    orc::fill_transform(dim, transform, xmin, xmax, c.transform, c.shift, &c.any_transform, tmin, tmax);
    return orc::integrate(orc::user_integrand, &c, dim, fdim, tmin, tmax, relTol, absTol, norm, maxEval);
}

int orc_status(void* r) { return ((orc::Result*)r)->status; }
long long orc_num_eval(void* r) { return ((orc::Result*)r)->numEval; }
int orc_converged(void* r) { return ((orc::Result*)r)->converged; }
long long orc_num_regions(void* r) { return (long long)((orc::Result*)r)->splitDim.size(); }
long long orc_num_iterations(void* r) { return (long long)((orc::Result*)r)->trace.batch.size(); }
void orc_get_result(void* r, double* out) {  // [2][fdim]: values then errors (Cubature.java:169)
    orc::Result* R = (orc::Result*)r;
    for (int j = 0; j < R->fdim; ++j) {
        out[j] = R->val[j];
        out[R->fdim + j] = R->err[j];
    }
}
void orc_get_batches(void* r, long long* out) {
    orc::Result* R = (orc::Result*)r;
    for (size_t i = 0; i < R->trace.batch.size(); ++i) out[i] = R->trace.batch[i];
}
void orc_get_regions(void* r, double* centers, double* halfwidths, int* splitDim, double* val, double* err,
                     double* errmax) {
    orc::Result* R = (orc::Result*)r;
    size_t n = R->splitDim.size();
    if (centers) std::memcpy(centers, R->centers.data(), sizeof(double) * n * R->dim);
    if (halfwidths) std::memcpy(halfwidths, R->halfwidths.data(), sizeof(double) * n * R->dim);
    if (splitDim) std::memcpy(splitDim, R->splitDim.data(), sizeof(int) * n);
    if (val) std::memcpy(val, R->rval.data(), sizeof(double) * n * R->fdim);
    if (err) std::memcpy(err, R->rerr.data(), sizeof(double) * n * R->fdim);
    if (errmax) std::memcpy(errmax, R->errmax.data(), sizeof(double) * n);
}
void orc_free(void* r) { delete (orc::Result*)r; }

// Rule evaluation alone on a caller-supplied list of boxes (parity tests of the device rule kernels).
// centers/halfwidths are [nR][dim] row-major IN t-SPACE; outputs val/err [nR][fdim], splitDim [nR].
int orc_eval_regions(int family, int dim, int fdim, const double* params, int nparams, const int* transform,
                     const double* shift, long long nR, const double* centers, const double* halfwidths, double* val,
                     double* err, int* splitDim) {
    orc::BuiltinCtx c;
    c.family = family;
    c.dim = dim;
    c.fdim = fdim;
    c.params.assign(params, params + nparams);
    c.use_libm = 0;
    c.any_transform = false;
    for (int d = 0; d < dim && d < 32; ++d) {
        c.transform[d] = transform ? transform[d] : 0;
        c.shift[d] = shift ? shift[d] : 0.0;
        if (c.transform[d]) c.any_transform = true;
    }
    orc::Rule* r = dim == 1 ? (orc::Rule*)new orc::RuleGaussKronrod_1d(dim, fdim) : (orc::Rule*)new orc::Rule75GenzMalik(dim, fdim);
    r->fn = orc::builtin_integrand;
    r->ctx = &c;
    std::vector<orc::Region*> R((size_t)nR);
    for (long long i = 0; i < nR; ++i) {
        orc::Hypercube h;
        h.dim = dim;
        h.centers.assign(centers + i * dim, centers + (i + 1) * dim);
        h.halfwidths.assign(halfwidths + i * dim, halfwidths + (i + 1) * dim);
        h.computeVolume();
        R[(size_t)i] = orc::Region::init(h, fdim);
    }
    r->evalError(R.data(), nR);
    for (long long i = 0; i < nR; ++i) {
        for (int j = 0; j < fdim; ++j) {
            val[i * fdim + j] = R[(size_t)i]->ee[j].val;
            err[i * fdim + j] = R[(size_t)i]->ee[j].err;
        }
        splitDim[i] = R[(size_t)i]->splitDim;
        delete R[(size_t)i];
    }
    delete r;
    return 0;
}

// Rule.converged alone (Rule.java:162-279) on dense vectors: 1 / 0, or -1 when both tolerances are NaN
int orc_converged_vec(int fdim, const double* val, const double* err, double absTol, double relTol, int norm) {
    std::vector<orc::ErrorEstimate> ee((size_t)fdim);
    for (int j = 0; j < fdim; ++j) {
        ee[(size_t)j].val = val[j];
        ee[(size_t)j].err = err[j];
    }
    return orc::converged(fdim, ee, absTol, relTol, norm);
}

// detmath entry points so tests can compare host bits with device bits and with mpmath
double orc_dm_exp(double x) { return dm_exp(x); }
double orc_dm_sin(double x) { return dm_sin(x); }
double orc_dm_cos(double x) { return dm_cos(x); }
double orc_dm_log(double x) { return dm_log(x); }
double orc_dm_pow(double x, double y) { return dm_pow(x, y); }
void orc_eval_family_points(int family, int dim, int fdim, const double* params, int nparams, long long n,
                            const double* x /*[dim][n]*/, double* out /*[fdim][n]*/) {
    orc::BuiltinCtx c;
    c.family = family;
    c.dim = dim;
    c.fdim = fdim;
    c.params.assign(params, params + nparams);
    c.use_libm = 0;
    c.any_transform = false;
    std::vector<const double*> xp((size_t)dim);
    std::vector<double*> fp((size_t)fdim);
    for (int d = 0; d < dim; ++d) xp[(size_t)d] = x + (size_t)d * n;
    for (int j = 0; j < fdim; ++j) fp[(size_t)j] = out + (size_t)j * n;
    orc::builtin_integrand(&c, dim, n, xp.data(), fdim, fp.data());
}

}  // extern "C"
