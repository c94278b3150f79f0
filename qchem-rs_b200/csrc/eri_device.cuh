// eri_device.cuh -- FP64 McMurchie-Davidson shell-quartet evaluation and J/K digestion for sm_100a.
//
// Replaces the arithmetic behind `molint::eri` (call sites core/src/hf/rhf.rs:45, uhf.rs:55; the crate
// is absent from the reference tree) fused with the density contraction of rhf.rs:152-167 and
// uhf.rs:210-227.  Everything is templated on the angular momenta so that every index below is a
// compile-time constant after unrolling: R_tuv, the E coefficients and the integral block live in
// registers, not in indexed local memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>

namespace qcf {

constexpr double PI_D = 3.14159265358979323846;
constexpr int LMAX = 2;
constexpr int BOYS_LTOT = 4 * LMAX;          // 8
constexpr double BOYS_TMAX = 64.0;          // extent of the table; each class switches at boys_tmax(L) <= 64
constexpr int BOYS_PER_UNIT = 16;            // grid step 1/16
constexpr int BOYS_NGRID = 64 * BOYS_PER_UNIT + 1;
constexpr int BOYS_ORDER = 6;                // Taylor order (7 terms), |dT| <= 1/32 -> 6e-15
constexpr int BOYS_ROW = 8;                  // doubles per grid point and class: 7 Taylor coefficients + exp(-T0)

// ---- compile-time index helpers ---------------------------------------------------------------
__host__ __device__ constexpr int ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int nherm(int l) { return (l + 1) * (l + 2) * (l + 3) / 6; }
// Cartesian component i of shell l: x,y,z / xx,xy,xz,yy,yz,zz
__host__ __device__ constexpr int cart_x(int l, int i) { return l == 0 ? 0 : l == 1 ? (i == 0) : (i == 0 ? 2 : (i <= 2 ? 1 : 0)); }
__host__ __device__ constexpr int cart_y(int l, int i) { return l == 0 ? 0 : l == 1 ? (i == 1) : (i == 3 ? 2 : ((i == 1 || i == 4) ? 1 : 0)); }
__host__ __device__ constexpr int cart_z(int l, int i) { return l - cart_x(l, i) - cart_y(l, i); }
__host__ __device__ constexpr int cart_pow(int l, int i, int axis) { return axis == 0 ? cart_x(l, i) : axis == 1 ? cart_y(l, i) : cart_z(l, i); }
// Hermite index (t,u,v), ordered by total order k = t+u+v
__host__ __device__ constexpr int hidx(int t, int u, int v) {
    return (t + u + v) * (t + u + v + 1) * (t + u + v + 2) / 6 + (u + v) * (u + v + 1) / 2 + v;
}

// ---- pair data layout --------------------------------------------------------------------------
// One "pair group" = all significant shell pairs of one (la >= lb) class with the same number of
// primitive pairs K, sorted by power-of-two Schwarz bucket (descending) and, inside a bucket, by shell
// indices, so that neighbouring lanes gather/scatter neighbouring density and Fock elements.
// Structure of arrays: field f of primitive k of pair i sits at prim[(k * PF_COUNT + f) * npair + i], so consecutive kets (= consecutive lanes)
// read consecutive doubles.  (Round 2 measured the alternatives on the N = 1007 build: 64-byte records per primitive read
// with 256-bit loads -- neutral for the block kernel, 20-40 % slower for the slab kernel; see profiles/README.md.)
enum PrimField { PF_P = 0, PF_PX, PF_PY, PF_PZ, PF_C, PF_PAX, PF_PAY, PF_PAZ, PF_COUNT };
struct PairGroup {
    int npair, K, la, lb;
    const int* fa;        // first basis function of shell a (the higher-l shell)
    const int* fb;
    const int* sa;        // shell ids (density-block screening)
    const int* sb;
    const int* nprim;     // primitive pairs kept for this pair (<= K; sorted by primitive Schwarz factor, negligible ones dropped)
    const double* Q;      // Schwarz factor
    const double* Qb;     // power-of-two bucket ceiling of Q, non-increasing along the list (prefix search)
    const double* prim;   // [K][PF_COUNT][npair];  PF_C = sqrt(2) pi^(5/4) c_a c_b exp(-mu AB^2) / p
    const double* AB;     // [3][npair]  A - B
    const float* Dp;      // [npair] max |P| over the pair's shell block, rewritten by every build (pair_dmax_kernel)
};

// Per-build scalars that live on the device, so that a build never has to wait for the host:
// the global density maximum (float bits, written by atomicMax) and the fixed-point scale of the deterministic mode.
struct BuildScalars {
    unsigned int dmax_bits;
    unsigned int pad;
    double fx_scale;          // 0: FP64 atomics (unordered);  > 0: accumulate round(v * fx_scale) with 64-bit integer atomics
    double abs_sum;           // sum |P| (fixed-order reduction), the bound the scale is derived from
};

struct BuildArgs {
    int N, nshell, nk;        // basis functions, shells, number of exchange densities (1 RHF, 2 UHF)
    const double* Pj;         // component-scaled Coulomb density (RHF: P, UHF: Pa+Pb), N x N
    const double* Pk0;        // exchange densities
    const double* Pk1;
    double* AJ;               // one-sided accumulators (symmetrised by the finalize kernel)
    double* AK0;
    double* AK1;
    const float* Dsh;         // [nshell][nshell] max |P| per shell block (over all densities)
    double tau;               // screening threshold (0: no screening)
    const BuildScalars* sc;   // device-resident per-build scalars (global density maximum, fixed-point scale)
    unsigned long long* counter;  // evaluated quartets of this launch
    const double* boys;       // [BOYS_LTOT+1][BOYS_NGRID][BOYS_ROW]
    const int* bra_list;      // this rank's share of the bra list (cost-balanced split); null: every bra pair
    double red_eps;           // scatter contributions below this magnitude are skipped (0.1 tau; 0 without screening)
    int ket_chunk;            // kets per CTA (grid.y strides over the ket list)
};

// ---- fast reciprocal square root / reciprocal (positive, normal arguments) --------------------------
// IEEE division and sqrt cost 20-40 instructions each in FP64 and were two thirds of all instructions
// executed by the (ss|ss) and (ps|ss) kernels.  MUFU.RSQ64H seed (2^-22) + two Newton steps: <= 2 ulp.
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    return y;
}
__device__ __forceinline__ double fast_rcp(double x) {
    const double r = fast_rsqrt(x);
    return r * r;
}

// ---- Boys function -----------------------------------------------------------------------------
// F_0..F_L(T).  T < boys_tmax(L) (36 .. 64): 7-term Taylor expansion of F_L about the nearest grid point (table row holds
// F_{L+k}(T0)/k!), then the stable downward recursion F_{m-1} = (2T F_m + e^-T)/(2m-1).  Above:
// F_0 = sqrt(pi/T)/2 and the upward recursion without the exp(-T) term (stable for T > m).
// Above boys_tmax(L) the asymptotic branch may drop exp(-T): exp(-T) (2T)^L / ((2L-1)!! F_0) < 1e-16 there.
__host__ __device__ constexpr double boys_tmax(int L) {
    return L == 0 ? 36.0 : L == 1 ? 44.0 : L == 2 ? 48.0 : L == 3 ? 52.0 : L == 4 ? 56.0 : L == 5 ? 58.0 : L == 6 ? 60.0 : L == 7 ? 62.0 : 64.0;
}
template <int L>
__device__ __forceinline__ void boys(double T, const double* __restrict__ table, double (&F)[L + 1]) {
    if (T < boys_tmax(L)) {
#ifdef QCF_FAKE_BOYS      // timing experiment only (wrong results): every lane reads table row 0
        const int g = 0;
#else
        const int g = (int)(T * BOYS_PER_UNIT + 0.5);
#endif
        const double d = (double)g * (1.0 / BOYS_PER_UNIT) - T;
        const double* r = table + ((size_t)L * BOYS_NGRID + g) * BOYS_ROW;
        double f = __ldg(r + 6);
#pragma unroll
        for (int k = 5; k >= 0; --k) f = fma(f, d, __ldg(r + k));
        F[L] = f;
        if constexpr (L > 0) {
            // exp(-T) = exp(-T0) exp(d), |d| <= 1/32: row slot 7 holds exp(-T0); 8-term Taylor (< 1e-19)
            double ed = 1.0 / 40320.0;
            ed = fma(ed, d, 1.0 / 5040.0); ed = fma(ed, d, 1.0 / 720.0); ed = fma(ed, d, 1.0 / 120.0);
            ed = fma(ed, d, 1.0 / 24.0); ed = fma(ed, d, 1.0 / 6.0); ed = fma(ed, d, 0.5);
            ed = fma(ed, d, 1.0); ed = fma(ed, d, 1.0);
            const double e = __ldg(r + 7) * ed;
            const double t2 = 2.0 * T;
#pragma unroll
            for (int m = L; m > 0; --m) F[m - 1] = fma(t2, F[m], e) * (1.0 / (2 * m - 1));
        }
    } else {
        // T >= boys_tmax(L): exp(-T) is below 1e-16 F_m(T) for every m <= L, so the upward recursion
        // F_{m+1} = ((2m+1) F_m - e^-T) / 2T loses nothing by dropping it -- no exp on this path at all
        const double rs = fast_rsqrt(T);
        F[0] = 0.88622692545275801365 * rs;       // sqrt(pi)/2 / sqrt(T)
        if constexpr (L > 0) {
            const double h = 0.5 * rs * rs;
#pragma unroll
            for (int m = 0; m < L; ++m) F[m + 1] = (2 * m + 1) * F[m] * h;
        }
    }
}

// ---- Hermite expansion coefficients for one axis ------------------------------------------------
// e[(i*(LB+1)+j)*(LA+LB+1)+t] = E_t^{ij}, E_0^{00} = 1 (the Gaussian prefactor travels in PF_C).
template <int LA, int LB>
struct EAxis {
    static constexpr int NT = LA + LB + 1;
    double e[(LA + 1) * (LB + 1) * NT];
    __device__ __forceinline__ double& at(int i, int j, int t) { return e[(i * (LB + 1) + j) * NT + t]; }
    __device__ __forceinline__ double get(int i, int j, int t) const { return e[(i * (LB + 1) + j) * NT + t]; }
    // sign = -1 folds the (-1)^t of the ket expansion into the table
    __device__ __forceinline__ void build(double h, double xpa, double xpb, bool ket_sign) {
        at(0, 0, 0) = 1.0;
#pragma unroll
        for (int i = 0; i <= LA; ++i) {
            if (i > 0) {
#pragma unroll
                for (int t = 0; t <= i; ++t) {
                    double v = (t <= i - 1) ? xpa * get(i - 1, 0, t) : 0.0;
                    if (t > 0) v = fma(h, get(i - 1, 0, t - 1), v);
                    if (t + 1 <= i - 1) v = fma((double)(t + 1), get(i - 1, 0, t + 1), v);
                    at(i, 0, t) = v;
                }
            }
#pragma unroll
            for (int j = 1; j <= LB; ++j) {
#pragma unroll
                for (int t = 0; t <= i + j; ++t) {
                    double v = (t <= i + j - 1) ? xpb * get(i, j - 1, t) : 0.0;
                    if (t > 0) v = fma(h, get(i, j - 1, t - 1), v);
                    if (t + 1 <= i + j - 1) v = fma((double)(t + 1), get(i, j - 1, t + 1), v);
                    at(i, j, t) = v;
                }
            }
        }
        if (ket_sign) {
#pragma unroll
            for (int i = 0; i <= LA; ++i)
#pragma unroll
                for (int j = 0; j <= LB; ++j)
#pragma unroll
                    for (int t = 1; t <= i + j; t += 2) at(i, j, t) = -get(i, j, t);
        }
    }
};

template <int LA, int LB>
struct PairE {
    EAxis<LA, LB> ax[3];
};

// ---- Hermite Coulomb integrals R_tuv, t+u+v <= L, in place ---------------------------------------
// On entry c[n] = pref * (-2 alpha)^n F_n.  Level n is built from level n+1, orders descending, so a
// single array of nherm(L) values suffices.
// Layout: compact (hidx) for the fully unrolled classes; for the large classes a dense (L+1)^3 cube so
// that R[p+q] = R[off(p) + off(q)] and the ket-transform loop over q can stay a runtime loop.
template <int L, bool CUBE>
__host__ __device__ constexpr int ridx(int t, int u, int v) {
    return CUBE ? (t * (L + 1) + u) * (L + 1) + v : hidx(t, u, v);
}
template <int L, bool CUBE>
__host__ __device__ constexpr int rsize() { return CUBE ? (L + 1) * (L + 1) * (L + 1) : nherm(L); }

template <int L, bool CUBE>
__device__ __forceinline__ void hermite_R(const double (&c)[L + 1], double X, double Y, double Z, double (&R)[rsize<L, CUBE>()]) {
#define hidx(t, u, v) ridx<L, CUBE>(t, u, v)
    R[0] = c[L];
#pragma unroll
    for (int n = L - 1; n >= 0; --n) {
#pragma unroll
        for (int k = L - n; k >= 1; --k) {
#pragma unroll
            for (int t = k; t >= 0; --t) {
#pragma unroll
                for (int u = k - t; u >= 0; --u) {
                    const int v = k - t - u;
                    double val;
                    if (t > 0) {
                        val = X * R[hidx(t - 1, u, v)];
                        if (t > 1) val = fma((double)(t - 1), R[hidx(t - 2, u, v)], val);
                    } else if (u > 0) {
                        val = Y * R[hidx(t, u - 1, v)];
                        if (u > 1) val = fma((double)(u - 1), R[hidx(t, u - 2, v)], val);
                    } else {
                        val = Z * R[hidx(t, u, v - 1)];
                        if (v > 1) val = fma((double)(v - 1), R[hidx(t, u, v - 2)], val);
                    }
                    R[hidx(t, u, v)] = val;
                }
            }
        }
        R[0] = c[n];
    }
#undef hidx
}

// Classes whose integral block does not fit the register file (LARGE) keep R (cube layout), the E
// tables and the block in local memory (L1) and run ONE out-of-line slab function with runtime ket
// component indices; all other classes are fully unrolled and inlined with compile-time indices.
// (The Fock build serves every LARGE class with the slab kernel of eri_slab.cuh; this path remains for
// the single-quartet parity kernel and the (dd|dd) Schwarz diagonal, which need the whole block at once.)
template <int LA, int LB, int LC, int LD>
struct ClassTraits {
    static constexpr int NI = ncart(LA) * ncart(LB) * ncart(LC) * ncart(LD);
    static constexpr bool LARGE = NI > 81;
};

// ---- one ket slab (ket components ic, id) of one primitive quartet --------------------------------
// I[(ia*NB+ib)*NCD + ic*ND+id] += sum_tuv E^ab_tuv sum_q (-1)^q E^cd_q R[tuv+q]
template <int LA, int LB, int LC, int LD>
__device__ __forceinline__ void ket_slab_impl(const double* __restrict__ R, const PairE<LA, LB>& Eab, const PairE<LC, LD>& Ecd,
                                              double* __restrict__ I, const int ic, const int id) {
    constexpr int L = LA + LB + LC + LD;
    constexpr bool CUBE = ClassTraits<LA, LB, LC, LD>::LARGE;
    constexpr int LAB = LA + LB;
    constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
    constexpr int NCD = NC * ND;
    const int cx = cart_x(LC, ic), cy = cart_y(LC, ic), cz = cart_z(LC, ic);
    const int dx = cart_x(LD, id), dy = cart_y(LD, id), dz = cart_z(LD, id);
    // ket Hermite -> Cartesian: H[tuv] = sum_{tau,nu,phi} Ecd R[t+tau,u+nu,v+phi]
    double H[nherm(LAB)];
#pragma unroll
    for (int i = 0; i < nherm(LAB); ++i) H[i] = 0.0;
#pragma unroll
    for (int tt = 0; tt <= cx + dx; ++tt) {
#pragma unroll
        for (int uu = 0; uu <= cy + dy; ++uu) {
#pragma unroll
            for (int vv = 0; vv <= cz + dz; ++vv) {
                const double w = Ecd.ax[0].get(cx, dx, tt) * Ecd.ax[1].get(cy, dy, uu) * Ecd.ax[2].get(cz, dz, vv);
                const double* __restrict__ Rq = R + ridx<L, CUBE>(tt, uu, vv);
#pragma unroll
                for (int t = 0; t <= LAB; ++t)
#pragma unroll
                    for (int u = 0; u <= LAB - t; ++u)
#pragma unroll
                        for (int v = 0; v <= LAB - t - u; ++v) {
                            const double r = CUBE ? Rq[ridx<L, true>(t, u, v)] : R[hidx(t + tt, u + uu, v + vv)];
                            H[hidx(t, u, v)] = fma(w, r, H[hidx(t, u, v)]);
                        }
            }
        }
    }
    // bra Hermite -> Cartesian
    double* __restrict__ Iout = I + ic * ND + id;
#pragma unroll
    for (int ia = 0; ia < NA; ++ia) {
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) {
            const int ax = cart_x(LA, ia), ay = cart_y(LA, ia), az = cart_z(LA, ia);
            const int bx = cart_x(LB, ib), by = cart_y(LB, ib), bz = cart_z(LB, ib);
            double s = 0.0;
#pragma unroll
            for (int t = 0; t <= ax + bx; ++t) {
#pragma unroll
                for (int u = 0; u <= ay + by; ++u) {
                    const double exy = Eab.ax[0].get(ax, bx, t) * Eab.ax[1].get(ay, by, u);
#pragma unroll
                    for (int v = 0; v <= az + bz; ++v)
                        s = fma(exy * Eab.ax[2].get(az, bz, v), H[hidx(t, u, v)], s);
                }
            }
            Iout[(ia * NB + ib) * NCD] += s;
        }
    }
}
template <int LA, int LB, int LC, int LD>
__device__ __noinline__ void ket_slab_call(const double* __restrict__ R, const PairE<LA, LB>& Eab, const PairE<LC, LD>& Ecd,
                                           double* __restrict__ I, int ic, int id) {
    ket_slab_impl<LA, LB, LC, LD>(R, Eab, Ecd, I, ic, id);
}

// ---- one primitive quartet: I[ab][cd] += [ab|cd] -------------------------------------------------
// cPcQ carries contraction coefficients, Gaussian prefactors and 2 pi^(5/2)/(pq); 1/sqrt(p+q) is applied here.
template <int LA, int LB, int LC, int LD>
__device__ __forceinline__ void prim_quartet(const PairE<LA, LB>& Eab, const PairE<LC, LD>& Ecd, double p, double q,
                                             double PQx, double PQy, double PQz, double cPcQ,
                                             const double* __restrict__ boys_table,
                                             double (&I)[ncart(LA) * ncart(LB) * ncart(LC) * ncart(LD)]) {
    constexpr int L = LA + LB + LC + LD;
    constexpr bool CUBE = ClassTraits<LA, LB, LC, LD>::LARGE;
    constexpr int NC = ncart(LC), ND = ncart(LD);
    const double pq = p + q;
    const double rspq = fast_rsqrt(pq);
    const double rpq = rspq * rspq;
    const double alpha = p * q * rpq;
    const double T = alpha * (PQx * PQx + PQy * PQy + PQz * PQz);
    double F[L + 1];
    boys<L>(T, boys_table, F);
    double c[L + 1];
    {
        double f = cPcQ * rspq;
        const double m2a = -2.0 * alpha;
#pragma unroll
        for (int n = 0; n <= L; ++n) { c[n] = f * F[n]; f *= m2a; }
    }
    double R[rsize<L, CUBE>()];
    hermite_R<L, CUBE>(c, PQx, PQy, PQz, R);
    if constexpr (CUBE) {
#pragma unroll 1
        for (int ic = 0; ic < NC; ++ic)
#pragma unroll 1
            for (int id = 0; id < ND; ++id) ket_slab_call<LA, LB, LC, LD>(R, Eab, Ecd, I, ic, id);
    } else {
#pragma unroll
        for (int ic = 0; ic < NC; ++ic)
#pragma unroll
            for (int id = 0; id < ND; ++id) ket_slab_impl<LA, LB, LC, LD>(R, Eab, Ecd, I, ic, id);
    }
}

// load primitive k of pair i and build its E tables
template <int LA, int LB>
__device__ __forceinline__ void load_prim(const PairGroup& g, int i, int k, double ABx, double ABy, double ABz,
                                          double& p, double& Px, double& Py, double& Pz, double& cP,
                                          PairE<LA, LB>& E, bool ket_sign) {
    const double* base = g.prim + (size_t)k * PF_COUNT * g.npair + i;
    const size_t np = g.npair;
    p = __ldg(base + PF_P * np);
    Px = __ldg(base + PF_PX * np);
    Py = __ldg(base + PF_PY * np);
    Pz = __ldg(base + PF_PZ * np);
    cP = __ldg(base + PF_C * np);
    if constexpr (LA + LB > 0) {
        const double h = 0.5 * fast_rcp(p);
        const double pax = __ldg(base + PF_PAX * np), pay = __ldg(base + PF_PAY * np), paz = __ldg(base + PF_PAZ * np);
        // P - B = (P - A) + (A - B)
        E.ax[0].build(h, pax, pax + ABx, ket_sign);
        E.ax[1].build(h, pay, pay + ABy, ket_sign);
        E.ax[2].build(h, paz, paz + ABz, ket_sign);
    } else {
        E.ax[0].e[0] = 1.0; E.ax[1].e[0] = 1.0; E.ax[2].e[0] = 1.0;
    }
}

// contracted quartet block of (bra pair ib_, ket pair ik_), scaled by `scale`
template <int LA, int LB, int LC, int LD>
__device__ __forceinline__ void contracted_quartet(const PairGroup& bra, int ib_, const PairGroup& ket, int ik_, double scale,
                                                   const double* __restrict__ boys_table,
                                                   double (&I)[ncart(LA) * ncart(LB) * ncart(LC) * ncart(LD)]) {
    constexpr int NI = ncart(LA) * ncart(LB) * ncart(LC) * ncart(LD);
#pragma unroll
    for (int i = 0; i < NI; ++i) I[i] = 0.0;
    double ABx = 0, ABy = 0, ABz = 0, CDx = 0, CDy = 0, CDz = 0;
    if constexpr (LB > 0) {
        ABx = __ldg(bra.AB + ib_); ABy = __ldg(bra.AB + bra.npair + ib_); ABz = __ldg(bra.AB + 2 * (size_t)bra.npair + ib_);
    }
    if constexpr (LD > 0) {
        CDx = __ldg(ket.AB + ik_); CDy = __ldg(ket.AB + ket.npair + ik_); CDz = __ldg(ket.AB + 2 * (size_t)ket.npair + ik_);
    }
    const int nkc = __ldg(ket.nprim + ik_), nkb = __ldg(bra.nprim + ib_);
    for (int kc = 0; kc < nkc; ++kc) {
        double q, Qx, Qy, Qz, cQ;
        PairE<LC, LD> Ecd;
        load_prim<LC, LD>(ket, ik_, kc, CDx, CDy, CDz, q, Qx, Qy, Qz, cQ, Ecd, true);
        cQ *= scale;
        for (int kb = 0; kb < nkb; ++kb) {
            double p, Px, Py, Pz, cP;
            PairE<LA, LB> Eab;
            load_prim<LA, LB>(bra, ib_, kb, ABx, ABy, ABz, p, Px, Py, Pz, cP, Eab, false);
            prim_quartet<LA, LB, LC, LD>(Eab, Ecd, p, q, Px - Qx, Py - Qy, Pz - Qz, cP * cQ, boys_table, I);
        }
    }
}

// ---- bra primitives staged in shared memory (block kernel) ------------------------------------------
// The bra pair is fixed per CTA: its primitive data (p, P, prefactor, P-A and 1/2p) is copied once into shared
// memory and read back as broadcast LDS -- no global address arithmetic and no reciprocal per primitive quartet.
constexpr int BRA_S = 9;    // doubles per staged bra primitive: PrimField order, then 1/(2p)
__device__ __forceinline__ void stage_bra(const PairGroup& g, int ib_, int nkb, double* __restrict__ dst) {
    for (int t = threadIdx.x; t < nkb * BRA_S; t += blockDim.x) {
        const int kb = t / BRA_S, f = t % BRA_S;
        const double* base = g.prim + (size_t)kb * PF_COUNT * g.npair + ib_;
        dst[t] = f < PF_COUNT ? __ldg(base + (size_t)f * g.npair) : 0.5 / __ldg(base + (size_t)PF_P * g.npair);
    }
}
template <int LA, int LB>
__device__ __forceinline__ void load_prim_staged(const double* __restrict__ b, double ABx, double ABy, double ABz,
                                                 double& p, double& Px, double& Py, double& Pz, double& cP, PairE<LA, LB>& E) {
    p = b[PF_P]; Px = b[PF_PX]; Py = b[PF_PY]; Pz = b[PF_PZ]; cP = b[PF_C];
    if constexpr (LA + LB > 0) {
        const double h = b[PF_COUNT];
        const double pax = b[PF_PAX], pay = b[PF_PAY], paz = b[PF_PAZ];
        E.ax[0].build(h, pax, pax + ABx, false);
        E.ax[1].build(h, pay, pay + ABy, false);
        E.ax[2].build(h, paz, paz + ABz, false);
    } else {
        E.ax[0].e[0] = 1.0; E.ax[1].e[0] = 1.0; E.ax[2].e[0] = 1.0;
    }
}
// contracted quartet block with the bra primitives read from the staged copy.
// PS > 1: PS consecutive lanes share this quartet; lane `sub` takes the primitive quartets sub, sub + PS, ... of the
// flattened (ket primitive, bra primitive) loop and the partial blocks are summed over the lane group afterwards
// (highly contracted classes: up to 36 x 36 primitive quartets per shell quartet would otherwise run serially in
// one thread while few such quartets exist -- a latency tail).
template <int LA, int LB, int LC, int LD, int PS>
__device__ __forceinline__ void contracted_quartet_staged(const double* __restrict__ bra_s, int nkb, double ABx, double ABy, double ABz,
                                                          const PairGroup& ket, int ik_, double scale,
                                                          const double* __restrict__ boys_table, int sub, unsigned int amask,
                                                          double (&I)[ncart(LA) * ncart(LB) * ncart(LC) * ncart(LD)]) {
    constexpr int NI = ncart(LA) * ncart(LB) * ncart(LC) * ncart(LD);
#pragma unroll
    for (int i = 0; i < NI; ++i) I[i] = 0.0;
    double CDx = 0, CDy = 0, CDz = 0;
    if constexpr (LD > 0) {
        CDx = __ldg(ket.AB + ik_); CDy = __ldg(ket.AB + ket.npair + ik_); CDz = __ldg(ket.AB + 2 * (size_t)ket.npair + ik_);
    }
    const int nkc = __ldg(ket.nprim + ik_);
    if constexpr (PS == 1) {
        for (int kc = 0; kc < nkc; ++kc) {
            double q, Qx, Qy, Qz, cQ;
            PairE<LC, LD> Ecd;
            load_prim<LC, LD>(ket, ik_, kc, CDx, CDy, CDz, q, Qx, Qy, Qz, cQ, Ecd, true);
            cQ *= scale;
            for (int kb = 0; kb < nkb; ++kb) {
                double p, Px, Py, Pz, cP;
                PairE<LA, LB> Eab;
                load_prim_staged<LA, LB>(bra_s + kb * BRA_S, ABx, ABy, ABz, p, Px, Py, Pz, cP, Eab);
                prim_quartet<LA, LB, LC, LD>(Eab, Ecd, p, q, Px - Qx, Py - Qy, Pz - Qz, cP * cQ, boys_table, I);
            }
        }
    } else {
        int kc = sub / nkb, kb = sub - kc * nkb;
        while (kc < nkc) {
            double q, Qx, Qy, Qz, cQ;
            PairE<LC, LD> Ecd;
            load_prim<LC, LD>(ket, ik_, kc, CDx, CDy, CDz, q, Qx, Qy, Qz, cQ, Ecd, true);
            cQ *= scale;
            for (; kb < nkb; kb += PS) {
                double p, Px, Py, Pz, cP;
                PairE<LA, LB> Eab;
                load_prim_staged<LA, LB>(bra_s + kb * BRA_S, ABx, ABy, ABz, p, Px, Py, Pz, cP, Eab);
                prim_quartet<LA, LB, LC, LD>(Eab, Ecd, p, q, Px - Qx, Py - Qy, Pz - Qz, cP * cQ, boys_table, I);
            }
            do { kb -= nkb; ++kc; } while (kb >= nkb);
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
#pragma unroll
            for (int o = PS / 2; o > 0; o >>= 1) I[i] += __shfl_xor_sync(amask, I[i], o);
        }
    }
}

// Accumulation into the global J / K matrices.  fx == 0: red.global.add.f64 (summation order = arrival order, results
// differ in the last bits from run to run).  fx > 0 (deterministic mode): the contribution is rounded to a multiple of
// 1/fx and added with a 64-bit INTEGER atomic -- integer addition is associative, so the result is bitwise independent
// of the arrival order; the finalize kernel converts back.
struct AccMode {
    double fx;      // fixed-point scale of the deterministic mode (0: FP64 atomics)
    double eps;     // contributions below this magnitude are not sent to memory (0: every contribution is)
};
__device__ __forceinline__ void red_add(double* addr, double v, const AccMode m) {
    // A shell quartet is evaluated when ONE of its six density-weighted bounds reaches tau; its other contributions can
    // be orders of magnitude smaller.  Each FP64 atomic costs ~1.3 SM-cycles per lane on the LSU path (the bottleneck of
    // the low-L classes), so contributions below eps = 0.1 tau -- a tenth of what the quartet screening itself neglects
    // per quartet -- are dropped.  Measured at N = 1007 (profiles/r2_ab_call7_red_threshold_factor.log): factor 0 77.7 ms,
    // 0.01 72.7 ms, 0.1 70.9 ms, 1 68.9 ms; max |dG| vs the unscreened oracle 3.5e-11, 3.5e-11, 3.7e-11, 2.1e-10.
    if (!(fabs(v) >= m.eps)) return;
    if (m.fx != 0.0) {
        const long long q = __double2ll_rn(v * m.fx);
        atomicAdd(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)q);
    } else {
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- digestion of one ket slab of a contracted block ---------------------------------------------
// Coulomb:  AJ[a,b] += P_cd I ;  AJ[c,d] += P_ab I
// Exchange: AK[a,c] += P_bd I ;  AK[a,d] += P_bc I ;  AK[b,c] += P_ad I ;  AK[b,d] += P_ac I
// kacc holds, per exchange density, [kac NA*NC | kad NA*ND | kbc NB*NC | kbd NB*ND].
template <int LA, int LB, int LC, int LD>
struct KAcc {
    static constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
    static constexpr int OFF_AC = 0, OFF_AD = NA * NC, OFF_BC = OFF_AD + NA * ND, OFF_BD = OFF_BC + NB * NC;
    static constexpr int SIZE = (NA + NB) * (NC + ND);
};

#ifdef QCF_FAKE_GATHER     // timing experiment only (wrong results): every density gather reads element 0
#define QCF_GIDX(x) ((size_t)0 * (x))
#else
#define QCF_GIDX(x) (x)
#endif
template <int LA, int LB, int LC, int LD, int NK>
__device__ __forceinline__ void digest_slab_impl(const double* __restrict__ I, double* __restrict__ jab,
                                                 const double* __restrict__ pab, const BuildArgs& a, const AccMode fx, int fa, int fb,
                                                 int fc, int fd, double* __restrict__ kacc, const int IC, const int ID) {
    using KA = KAcc<LA, LB, LC, LD>;
    constexpr int NA = KA::NA, NB = KA::NB, NC = KA::NC, ND = KA::ND;
    constexpr int NAB = NA * NB, NCD = NC * ND;
    const int ICD = IC * ND + ID;
    const int N = a.N;
    {
        const double pcd = __ldg(a.Pj + QCF_GIDX((size_t)(fc + IC) * N + fd + ID));
        double s = 0.0;
#pragma unroll
        for (int iab = 0; iab < NAB; ++iab) {
            const double v = I[iab * NCD + ICD];
            jab[iab] = fma(v, pcd, jab[iab]);
            s = fma(v, pab[iab], s);
        }
        red_add(a.AJ + (size_t)(fc + IC) * N + fd + ID, s, fx);
    }
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
        const double* __restrict__ Pk = kk == 0 ? a.Pk0 : a.Pk1;
        double* __restrict__ acc = kacc + kk * KA::SIZE;
        double pbd[NB], pbc[NB], pad[NA], pac[NA];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            pbd[i] = __ldg(Pk + QCF_GIDX((size_t)(fb + i) * N + fd + ID));
            pbc[i] = __ldg(Pk + QCF_GIDX((size_t)(fb + i) * N + fc + IC));
        }
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            pad[i] = __ldg(Pk + QCF_GIDX((size_t)(fa + i) * N + fd + ID));
            pac[i] = __ldg(Pk + QCF_GIDX((size_t)(fa + i) * N + fc + IC));
        }
        double kbc_s[NB], kbd_s[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) { kbc_s[i] = 0.0; kbd_s[i] = 0.0; }
#pragma unroll
        for (int ia = 0; ia < NA; ++ia) {
            double kac_s = 0.0, kad_s = 0.0;
#pragma unroll
            for (int ib = 0; ib < NB; ++ib) {
                const double v = I[(ia * NB + ib) * NCD + ICD];
                kac_s = fma(v, pbd[ib], kac_s);
                kad_s = fma(v, pbc[ib], kad_s);
                kbc_s[ib] = fma(v, pad[ia], kbc_s[ib]);
                kbd_s[ib] = fma(v, pac[ia], kbd_s[ib]);
            }
            acc[KA::OFF_AC + ia * NC + IC] += kac_s;
            acc[KA::OFF_AD + ia * ND + ID] += kad_s;
        }
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) {
            acc[KA::OFF_BC + ib * NC + IC] += kbc_s[ib];
            acc[KA::OFF_BD + ib * ND + ID] += kbd_s[ib];
        }
    }
}
template <int LA, int LB, int LC, int LD, int NK>
__device__ __noinline__ void digest_slab_call(const double* __restrict__ I, double* __restrict__ jab,
                                              const double* __restrict__ pab, const BuildArgs& a, const AccMode fx, int fa, int fb, int fc,
                                              int fd, double* __restrict__ kacc, int ic, int id) {
    digest_slab_impl<LA, LB, LC, LD, NK>(I, jab, pab, a, fx, fa, fb, fc, fd, kacc, ic, id);
}
template <int LA, int LB, int LC, int LD, int NK>
__device__ __forceinline__ void digest_all(const double* __restrict__ I, double* __restrict__ jab, const double* __restrict__ pab,
                                           const BuildArgs& a, const AccMode fx, int fa, int fb, int fc, int fd, double* __restrict__ kacc) {
    constexpr int NC = ncart(LC), ND = ncart(LD);
    if constexpr (ClassTraits<LA, LB, LC, LD>::LARGE) {
#pragma unroll 1
        for (int ic = 0; ic < NC; ++ic)
#pragma unroll 1
            for (int id = 0; id < ND; ++id) digest_slab_call<LA, LB, LC, LD, NK>(I, jab, pab, a, fx, fa, fb, fc, fd, kacc, ic, id);
    } else {
#pragma unroll
        for (int ic = 0; ic < NC; ++ic)
#pragma unroll
            for (int id = 0; id < ND; ++id) digest_slab_impl<LA, LB, LC, LD, NK>(I, jab, pab, a, fx, fa, fb, fc, fd, kacc, ic, id);
    }
}

// ---- ket scan shared by the two Fock-build kernels -----------------------------------------------------
// One scan step: every lane tests SW candidate kets (scan + j*32 + lane), all loads issued up front (coalesced:
// Q, the pair's own density maximum Dp, the two shell ids), then the density-weighted screening
//     Q_ab Q_cd max(D_ab, D_cd, max(D_ac, D_ad, D_bc, D_bd)/2) >= tau,
// and the survivors are appended to the warp's queue in list order.  Returns the new queue length.
// The scan is latency-bound (two dependent memory round trips per step), so its throughput is the number of
// candidates in flight: SW = 4 keeps 128 per warp instead of 32.
template <int SW>
__device__ __forceinline__ int scan_kets(const PairGroup& ket, const double tau, const int scan, const int nket, const double qab,
                                         const float dab, const float* __restrict__ dsh_a, const float* __restrict__ dsh_b,
                                         int* __restrict__ queue, int qn, const int lane) {
    bool ok[SW];
    if (tau > 0.0) {
        double qcd[SW];
        float dcd[SW];
        int sc[SW], sd[SW];
#pragma unroll
        for (int j = 0; j < SW; ++j) {
            const int ikc = scan + j * 32 + lane;
            ok[j] = ikc < nket;
            const int ii = ok[j] ? ikc : scan;          // scan < nket: always a valid index
            qcd[j] = __ldg(ket.Q + ii); dcd[j] = __ldg(ket.Dp + ii);
            sc[j] = __ldg(ket.sa + ii); sd[j] = __ldg(ket.sb + ii);
        }
#pragma unroll
        for (int j = 0; j < SW; ++j) {
            float dm = fmaxf(dab, dcd[j]);
            const float dk = fmaxf(fmaxf(dsh_a[sc[j]], dsh_a[sd[j]]), fmaxf(dsh_b[sc[j]], dsh_b[sd[j]]));
            dm = fmaxf(dm, 0.5f * dk);
            ok[j] = ok[j] && !(qab * qcd[j] * (double)dm < tau);
        }
    } else {
#pragma unroll
        for (int j = 0; j < SW; ++j) ok[j] = scan + j * 32 + lane < nket;
    }
#pragma unroll
    for (int j = 0; j < SW; ++j) {
        const unsigned int m = __ballot_sync(0xffffffffu, ok[j]);
        if (ok[j]) queue[qn + __popc(m & ((1u << lane) - 1u))] = scan + j * 32 + lane;
        qn += __popc(m);
    }
    return qn;
}

// ket prefix of bra pair ib_ that can survive Q_ab Q_cd Dmax >= tau (ket list sorted by descending Q bucket),
// clipped to this CTA's chunk [ket0, ket0 + chunk); returns the end index (<= ket0: nothing to do)
__device__ __forceinline__ int ket_prefix_end(const PairGroup& ket, const BuildArgs& a, double qab, int ib_, int same_group, int ket0) {
    int nket = ket.npair;
    if (a.tau > 0.0) {
        const double dmax = fmax((double)__uint_as_float(a.sc->dmax_bits), 1e-300);
        const double need = a.tau / (qab * dmax);
        int lo = 0, hi = ket.npair;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(ket.Qb + mid) >= need) lo = mid + 1; else hi = mid; }
        nket = lo;
    }
    if (same_group && nket > ib_ + 1) nket = ib_ + 1;
    if (nket > ket0 + a.ket_chunk) nket = ket0 + a.ket_chunk;
    return nket;
}

// ---- the block Fock-build kernel: one CTA per (bra pair, chunk of its ket prefix); s/p/ds bras -------
// NK = number of exchange densities (1: RHF / single-density J,K;  2: UHF alpha,beta).
// PS = lanes per shell quartet (1, or 4 / 8 for the highly contracted launches, see contracted_quartet_staged).
#ifndef QCF_SCANW
#define QCF_SCANW 4
#endif
template <int PS> struct BlockCfg {
    static constexpr int SW = PS == 1 ? QCF_SCANW : 1;        // scan width: a PS > 1 warp consumes only 32/PS quartets per pass
    static constexpr int QLEN = 32 * (SW + 1);
};
// WIDE: scan steps of SW candidates per lane; chosen by the launcher when the chunk gives every warp a full wide step.
// Short chunks (small molecules, a rank's share of few bras: kets per CTA are cut fine to fill the GPU) use the
// instantiation that scans 32 candidates per step, so that no warp idles.
template <int LA, int LB, int LC, int LD, int NK, int PS, bool WIDE>
__global__ void __launch_bounds__(128)
eri_jk_kernel(PairGroup bra, PairGroup ket, BuildArgs a, int same_group) {
    using KA = KAcc<LA, LB, LC, LD>;
    constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
    constexpr int NAB = NA * NB, NCD = NC * ND, NI = NAB * NCD;
    constexpr int SW = WIDE ? BlockCfg<PS>::SW : 1, QLEN = BlockCfg<PS>::QLEN, QPW = 32 / PS;
    const int ib_ = a.bra_list ? __ldg(a.bra_list + blockIdx.x) : (int)blockIdx.x;
    const double qab = __ldg(bra.Q + ib_);
    const int ket0 = blockIdx.y * a.ket_chunk;
    const int nket = ket_prefix_end(ket, a, qab, ib_, same_group, ket0);
    if (nket <= ket0) return;
    const AccMode fx{a.sc->fx_scale, a.red_eps};

    const int N = a.N;
    const int fa = __ldg(bra.fa + ib_), fb = __ldg(bra.fb + ib_);
    const int sa = __ldg(bra.sa + ib_), sb = __ldg(bra.sb + ib_);
    const double bra_deg = (sa == sb) ? 0.5 : 1.0;
    const float dab = __ldg(bra.Dp + ib_);

    // CTA-constant data in shared memory: the bra primitives and the two rows (shell a, shell b) of the
    // shell-block density maxima that the exchange screening gathers from
    extern __shared__ __align__(16) double dyn_s[];
    const int nkb = __ldg(bra.nprim + ib_);
    double* const bra_s = dyn_s;                                                     // [bra.K][BRA_S]
    float* const dsh_a = reinterpret_cast<float*>(dyn_s + (size_t)bra.K * BRA_S);    // [nshell]
    float* const dsh_b = dsh_a + a.nshell;                                           // [nshell]
    stage_bra(bra, ib_, nkb, bra_s);
    if (a.tau > 0.0)
        for (int t = threadIdx.x; t < a.nshell; t += blockDim.x) {
            dsh_a[t] = __ldg(a.Dsh + (size_t)sa * a.nshell + t);
            dsh_b[t] = __ldg(a.Dsh + (size_t)sb * a.nshell + t);
        }
    double ABx = 0, ABy = 0, ABz = 0;
    if constexpr (LB > 0) {
        ABx = __ldg(bra.AB + ib_); ABy = __ldg(bra.AB + bra.npair + ib_); ABz = __ldg(bra.AB + 2 * (size_t)bra.npair + ib_);
    }
    __syncthreads();

    double jab[NAB];
    double pab[NAB];
#pragma unroll
    for (int i = 0; i < NAB; ++i) { jab[i] = 0.0; pab[i] = __ldg(a.Pj + (size_t)(fa + i / NB) * N + fb + i % NB); }
    unsigned int nq = 0;

    // Warp-level compaction of the surviving kets: every warp scans 32*SW candidate kets at a time, applies the
    // density-weighted screening, and queues the survivors in shared memory; the expensive part below always
    // runs on (up to) 32/PS survivors, PS lanes each, so screened-out kets cost a scan step instead of an idle
    // lane for a whole contracted quartet.
    __shared__ int ket_queue[4][QLEN];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane / PS, sub = lane % PS;
    int* const queue = ket_queue[warp];
    int qn = 0;
    int scan = ket0 + warp * (32 * SW);
    const int scan_stride = (int)blockDim.x * SW;
    while (true) {
        while (qn < QPW && scan < nket) {
            qn = scan_kets<SW>(ket, a.tau, scan, nket, qab, dab, dsh_a, dsh_b, queue, qn, lane);
            scan += scan_stride;
        }
        __syncwarp();
        const int nrun = qn < QPW ? qn : QPW;
        if (nrun == 0) break;
        const unsigned int amask = __ballot_sync(0xffffffffu, slot < nrun);
        if (slot < nrun) {
        const int ik_ = queue[qn - nrun + slot];
        const int sc = __ldg(ket.sa + ik_), sd = __ldg(ket.sb + ik_);
        if (sub == 0) ++nq;
        const int fc = __ldg(ket.fa + ik_), fd = __ldg(ket.fb + ik_);
        double deg = bra_deg * ((sc == sd) ? 0.5 : 1.0);
        if (same_group && ik_ == ib_) deg *= 0.5;

        double I[NI];
        contracted_quartet_staged<LA, LB, LC, LD, PS>(bra_s, nkb, ABx, ABy, ABz, ket, ik_, deg, a.boys, sub, amask, I);

        if (sub == 0) {
        double kacc[NK * KA::SIZE];
#pragma unroll
        for (int i = 0; i < NK * KA::SIZE; ++i) kacc[i] = 0.0;
        digest_all<LA, LB, LC, LD, NK>(I, jab, pab, a, fx, fa, fb, fc, fd, kacc);
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) {
            double* __restrict__ AK = kk == 0 ? a.AK0 : a.AK1;
            const double* acc = kacc + kk * KA::SIZE;
#pragma unroll
            for (int i = 0; i < NA; ++i) {
#pragma unroll
                for (int j = 0; j < NC; ++j) red_add(AK + (size_t)(fa + i) * N + fc + j, acc[KA::OFF_AC + i * NC + j], fx);
#pragma unroll
                for (int j = 0; j < ND; ++j) red_add(AK + (size_t)(fa + i) * N + fd + j, acc[KA::OFF_AD + i * ND + j], fx);
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
#pragma unroll
                for (int j = 0; j < NC; ++j) red_add(AK + (size_t)(fb + i) * N + fc + j, acc[KA::OFF_BC + i * NC + j], fx);
#pragma unroll
                for (int j = 0; j < ND; ++j) red_add(AK + (size_t)(fb + i) * N + fd + j, acc[KA::OFF_BD + i * ND + j], fx);
            }
        }
        }   // sub == 0
        }   // slot < nrun
        qn -= nrun;
        __syncwarp();
    }

    // ---- J_ab: reduce over the CTA (fixed order), one atomic per element ----
    __shared__ double red[4][NAB];
#pragma unroll
    for (int i = 0; i < NAB; ++i) {
        const double s = warp_sum(jab[i]);
        if (lane == 0) red[warp][i] = s;
    }
    nq = __reduce_add_sync(0xffffffffu, nq);
    __shared__ unsigned int nqs[4];
    if (lane == 0) nqs[warp] = nq;
    __syncthreads();
    const int nwarp = (blockDim.x + 31) >> 5;
    for (int i = threadIdx.x; i < NAB; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += red[w][i];
        red_add(a.AJ + (size_t)(fa + i / NB) * N + fb + i % NB, s, fx);
    }
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int w = 0; w < nwarp; ++w) t += nqs[w];
        if (t) atomicAdd(a.counter, (unsigned long long)t);
    }
}

// ---- Schwarz factors: one thread per pair, Q = sqrt(max_ij |(ij|ij)|) ----------------------------
template <int LA, int LB>
__global__ void schwarz_kernel(PairGroup g, const double* __restrict__ boys_table, double* __restrict__ Qout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.npair) return;
    constexpr int NAB = ncart(LA) * ncart(LB);
    double I[NAB * NAB];
    contracted_quartet<LA, LB, LA, LB>(g, i, g, i, 1.0, boys_table, I);
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < NAB; ++k) m = fmax(m, fabs(I[k * NAB + k]));
    Qout[i] = sqrt(m);
}

// ---- a single contracted quartet written out (parity tests) --------------------------------------
template <int LA, int LB, int LC, int LD>
__global__ void quartet_kernel(PairGroup bra, int ib_, PairGroup ket, int ik_, const double* __restrict__ boys_table,
                               double* __restrict__ out) {
    constexpr int NI = ncart(LA) * ncart(LB) * ncart(LC) * ncart(LD);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double I[NI];
    contracted_quartet<LA, LB, LC, LD>(bra, ib_, ket, ik_, 1.0, boys_table, I);
#pragma unroll
    for (int i = 0; i < NI; ++i) out[i] = I[i];
}

// ---- launch interface of one angular class (defined in eri_class.cu, one object per class) ------
struct ClassLaunch {
    // ps = lanes per shell quartet (1, 4 or 8; block kernel only)
    void (*jk)(int nk, int ps, int nbra, int nket_max, int block, int kets_per_thread, cudaStream_t s, const PairGroup& bra, const PairGroup& ket, const BuildArgs& a, int same);
    // one-off per device: opt in to the dynamic shared memory the largest launch of this class needs
    cudaError_t (*init)(int max_bra_K, int nshell);
    int max_ps;           // largest ps this class was compiled for
    void (*quartet)(cudaStream_t s, const PairGroup& bra, int ib_, const PairGroup& ket, int ik_, const double* boys, double* out);
    void (*schwarz)(int grid, int block, cudaStream_t s, const PairGroup& g, const double* boys, double* Q);  // null unless (LA,LB)==(LC,LD)
};

}  // namespace qcf
