// eri_slab.cuh -- register-resident "slab" Fock-build kernel for the classes whose integral block does not
// fit the register file (all d-bra classes).
//
// Same arithmetic as eri_device.cuh (McMurchie-Davidson, behind molint::eri, rhf.rs:45 / uhf.rs:55, fused
// with the contractions of rhf.rs:152-167 / uhf.rs:210-227), different work decomposition:
//   * one CTA = one bra pair x one chunk of the ket list.  Everything that depends on the bra pair only --
//     primitive data and the combined Hermite coefficients E3[ab][tuv] = Ex_t Ey_u Ez_v -- is built once per
//     CTA in shared memory and then read as warp-uniform (broadcast) operands;
//   * one thread = one ket pair x one group of SPT ket components ("slabs"); the slab group is CTA-uniform
//     (grid.z), so each CTA runs straight-line code specialised for its ket components with every index a compile-time
//     constant (R_tuv, the ket-contracted Hermite integrals Hsum and the accumulators stay in registers);
//   * the ket primitives are contracted at the Hermite level, Hsum[tuv] = sum_kc sum_q E^cd_q R_{tuv+q}, so the
//     bra transform and the digestion run once per (quartet, bra primitive), never per primitive quartet;
//   * the integral block is never stored: every (ab|cd) value is digested into J and K as it is produced.
#pragma once
#include "eri_device.cuh"

namespace qcf {

// ---- layout of the combined bra coefficients ---------------------------------------------------------
template <int LA, int LB>
struct E3Layout {
    static constexpr int NA = ncart(LA), NB = ncart(LB), NAB = NA * NB;
    __host__ __device__ static constexpr int bx(int iab, int axis) {
        return cart_pow(LA, iab / NB, axis) + cart_pow(LB, iab % NB, axis) + 1;
    }
    __host__ __device__ static constexpr int box(int iab) { return bx(iab, 0) * bx(iab, 1) * bx(iab, 2); }
    __host__ __device__ static constexpr int off(int iab) {   // even offsets: 16-byte aligned pairs
        int o = 0;
        for (int j = 0; j < iab; ++j) o += (box(j) + 1) & ~1;
        return o;
    }
    static constexpr int SIZE = off(NAB);
    static constexpr int EA_N = (LA + 1) * (LB + 1) * (LA + LB + 1);   // entries of one EAxis table
    static constexpr int PRIM = 8;                                     // p, Px, Py, Pz, cP, pad
    static constexpr int E3_OFF = PRIM + ((3 * EA_N + 1) & ~1);        // even: E3 pairs are 16-byte aligned
    static constexpr int STRIDE = E3_OFF + SIZE;                       // doubles per bra primitive (even)
};

template <int LA, int LB, int LC, int LD, int NK, int SPT>
struct SlabCfg {
    static constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
    static constexpr int NAB = NA * NB, NCD = NC * ND;
    static_assert(NCD % SPT == 0, "slabs per thread must divide the number of ket components");
    static constexpr int G = NCD / SPT;                                // slab groups (grid.z; CTA-uniform)
    static constexpr int NSUB = 4;                                     // warps per CTA, each 32 kets per pass
    static constexpr int BLOCK = 32 * NSUB;
    static constexpr bool JSMEM = NAB > 18;                            // J_ab accumulators in shared memory
    static constexpr int LAB = LA + LB, L = LA + LB + LC + LD;
    static constexpr int NH = nherm(LAB);
};

// ket Hermite -> Cartesian for one ket component, accumulated:  H[p] += sum_q E^cd_q R[p+q]
template <int LA, int LB, int LC, int LD, int ICD>
__device__ __forceinline__ void ket_accumulate(const double (&R)[nherm(LA + LB + LC + LD)], const PairE<LC, LD>& Ecd,
                                               double (&H)[nherm(LA + LB)]) {
    constexpr int LAB = LA + LB;
    constexpr int ND = ncart(LD);
    constexpr int ic = ICD / ND, id = ICD % ND;
    constexpr int cx = cart_x(LC, ic), cy = cart_y(LC, ic), cz = cart_z(LC, ic);
    constexpr int dx = cart_x(LD, id), dy = cart_y(LD, id), dz = cart_z(LD, id);
#pragma unroll
    for (int tt = 0; tt <= cx + dx; ++tt)
#pragma unroll
        for (int uu = 0; uu <= cy + dy; ++uu)
#pragma unroll
            for (int vv = 0; vv <= cz + dz; ++vv) {
                const double w = Ecd.ax[0].get(cx, dx, tt) * Ecd.ax[1].get(cy, dy, uu) * Ecd.ax[2].get(cz, dz, vv);
#pragma unroll
                for (int t = 0; t <= LAB; ++t)
#pragma unroll
                    for (int u = 0; u <= LAB - t; ++u)
#pragma unroll
                        for (int v = 0; v <= LAB - t - u; ++v)
                            H[hidx(t, u, v)] = fma(w, R[hidx(t + tt, u + uu, v + vv)], H[hidx(t, u, v)]);
            }
}

// ---- bra transform + digestion of one slab group for one bra primitive -------------------------------
// Every index is a template parameter (IA, IB, the slab group SG), so box sizes, E3 offsets and the
// Hermite indices are constant expressions: Hsum stays in registers and no index arithmetic is executed.
template <int NK, int SPT, int NB>
struct DigestState {
    double pcd[SPT], scd[SPT];
    double pbd[NK][SPT][NB], pbc[NK][SPT][NB], kbc[NK][SPT][NB], kbd[NK][SPT][NB];
    double pad[NK][SPT], pac[NK][SPT], kac[NK][SPT], kad[NK][SPT];
};

template <int LA, int LB, int SPT, int IAB>
__device__ __forceinline__ void bra_value(const double (&Hs)[SPT][nherm(LA + LB)], const double* __restrict__ e3, double (&v)[SPT]) {
    using EL = E3Layout<LA, LB>;
    constexpr int nby = EL::bx(IAB, 1), nbz = EL::bx(IAB, 2), nbox = EL::box(IAB), off = EL::off(IAB);
#pragma unroll
    for (int s = 0; s < SPT; ++s) v[s] = 0.0;
#pragma unroll
    for (int k2 = 0; k2 < nbox / 2; ++k2) {
        const double2 ee = *reinterpret_cast<const double2*>(e3 + off + 2 * k2);
        {
            const int k = 2 * k2;
            const int t = k / (nby * nbz), u = (k / nbz) % nby, w = k % nbz;
#pragma unroll
            for (int s = 0; s < SPT; ++s) v[s] = fma(ee.x, Hs[s][hidx(t, u, w)], v[s]);
        }
        {
            const int k = 2 * k2 + 1;
            const int t = k / (nby * nbz), u = (k / nbz) % nby, w = k % nbz;
#pragma unroll
            for (int s = 0; s < SPT; ++s) v[s] = fma(ee.y, Hs[s][hidx(t, u, w)], v[s]);
        }
    }
    if constexpr (nbox & 1) {
        constexpr int k = nbox - 1;
        constexpr int t = k / (nby * nbz), u = (k / nbz) % nby, w = k % nbz;
        const double e = e3[off + k];
#pragma unroll
        for (int s = 0; s < SPT; ++s) v[s] = fma(e, Hs[s][hidx(t, u, w)], v[s]);
    }
}

template <int LA, int LB, int LC, int LD, int NK, int SPT, int IA, int IB>
__device__ __forceinline__ void digest_ab(const double (&Hs)[SPT][nherm(LA + LB)], const double* __restrict__ e3,
                                          const double* __restrict__ pab_s, double* __restrict__ jab_s,
                                          double (&jab)[ncart(LA) * ncart(LB)], DigestState<NK, SPT, ncart(LB)>& d) {
    using C = SlabCfg<LA, LB, LC, LD, NK, SPT>;
    constexpr int NB = C::NB, iab = IA * NB + IB;
    double v[SPT];
    bra_value<LA, LB, SPT, iab>(Hs, e3, v);
    const double pab = pab_s[iab];
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        if constexpr (C::JSMEM) jab_s[iab * C::BLOCK] = fma(v[s], d.pcd[s], jab_s[iab * C::BLOCK]);
        else jab[iab] = fma(v[s], d.pcd[s], jab[iab]);
        d.scd[s] = fma(v[s], pab, d.scd[s]);
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) {
            d.kac[kk][s] = fma(v[s], d.pbd[kk][s][IB], d.kac[kk][s]);
            d.kad[kk][s] = fma(v[s], d.pbc[kk][s][IB], d.kad[kk][s]);
            d.kbc[kk][s][IB] = fma(v[s], d.pad[kk][s], d.kbc[kk][s][IB]);
            d.kbd[kk][s][IB] = fma(v[s], d.pac[kk][s], d.kbd[kk][s][IB]);
        }
    }
}

template <int LA, int LB, int LC, int LD, int NK, int SPT, int IA, int... IB>
__device__ __forceinline__ void digest_a(std::integer_sequence<int, IB...>, const int SG, const double (&Hs)[SPT][nherm(LA + LB)],
                                         const double* __restrict__ e3, const double* __restrict__ pab_s,
                                         double* __restrict__ jab_s, double (&jab)[ncart(LA) * ncart(LB)],
                                         DigestState<NK, SPT, ncart(LB)>& d, const BuildArgs& a, const AccMode fx, int fa, int fc, int fd) {
    constexpr int ND = ncart(LD);
    const int N = a.N;
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int icd = SG * SPT + s, IC = icd / ND, ID = icd % ND;
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) {
            const double* __restrict__ Pk = kk == 0 ? a.Pk0 : a.Pk1;
            d.pad[kk][s] = __ldg(Pk + (size_t)(fa + IA) * N + fd + ID);
            d.pac[kk][s] = __ldg(Pk + (size_t)(fa + IA) * N + fc + IC);
            d.kac[kk][s] = 0.0; d.kad[kk][s] = 0.0;
        }
    }
    (digest_ab<LA, LB, LC, LD, NK, SPT, IA, IB>(Hs, e3, pab_s, jab_s, jab, d), ...);
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int icd = SG * SPT + s, IC = icd / ND, ID = icd % ND;
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) {
            double* __restrict__ AK = kk == 0 ? a.AK0 : a.AK1;
            red_add(AK + (size_t)(fa + IA) * N + fc + IC, d.kac[kk][s], fx);
            red_add(AK + (size_t)(fa + IA) * N + fd + ID, d.kad[kk][s], fx);
        }
    }
}

template <int LA, int LB, int LC, int LD, int NK, int SPT, int... IA>
__device__ __forceinline__ void digest_all_a(std::integer_sequence<int, IA...>, const int SG, const double (&Hs)[SPT][nherm(LA + LB)],
                                             const double* __restrict__ e3, const double* __restrict__ pab_s,
                                             double* __restrict__ jab_s, double (&jab)[ncart(LA) * ncart(LB)],
                                             DigestState<NK, SPT, ncart(LB)>& d, const BuildArgs& a, const AccMode fx, int fa, int fc, int fd) {
    (digest_a<LA, LB, LC, LD, NK, SPT, IA>(std::make_integer_sequence<int, ncart(LB)>{}, SG, Hs, e3, pab_s, jab_s, jab, d, a, fx, fa, fc, fd), ...);
}

// The slab group enters only through the ket component offsets (IC, ID) of the global addresses, so it is a
// runtime value here: one copy of the bra transform serves every slab group (code size, compile time).
template <int LA, int LB, int LC, int LD, int NK, int SPT>
__device__ __forceinline__ void slab_digest(const int SG, const double (&Hs)[SPT][nherm(LA + LB)], const double* __restrict__ e3,
                                            const double* __restrict__ pab_s, double* __restrict__ jab_s,
                                            double (&jab)[ncart(LA) * ncart(LB)], const BuildArgs& a, const AccMode fx, int fa, int fb, int fc, int fd) {
    using C = SlabCfg<LA, LB, LC, LD, NK, SPT>;
    constexpr int NA = C::NA, NB = C::NB, ND = C::ND;
    const int N = a.N;
    DigestState<NK, SPT, NB> d;
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int icd = SG * SPT + s, IC = icd / ND, ID = icd % ND;
        d.pcd[s] = __ldg(a.Pj + (size_t)(fc + IC) * N + fd + ID);
        d.scd[s] = 0.0;
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) {
            const double* __restrict__ Pk = kk == 0 ? a.Pk0 : a.Pk1;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                d.pbd[kk][s][i] = __ldg(Pk + (size_t)(fb + i) * N + fd + ID);
                d.pbc[kk][s][i] = __ldg(Pk + (size_t)(fb + i) * N + fc + IC);
                d.kbc[kk][s][i] = 0.0; d.kbd[kk][s][i] = 0.0;
            }
        }
    }
    digest_all_a<LA, LB, LC, LD, NK, SPT>(std::make_integer_sequence<int, NA>{}, SG, Hs, e3, pab_s, jab_s, jab, d, a, fx, fa, fc, fd);
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int icd = SG * SPT + s, IC = icd / ND, ID = icd % ND;
        red_add(a.AJ + (size_t)(fc + IC) * N + fd + ID, d.scd[s], fx);
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) {
            double* __restrict__ AK = kk == 0 ? a.AK0 : a.AK1;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                red_add(AK + (size_t)(fb + i) * N + fc + IC, d.kbc[kk][s][i], fx);
                red_add(AK + (size_t)(fb + i) * N + fd + ID, d.kbd[kk][s][i], fx);
            }
        }
    }
}

template <int LA, int LB, int LC, int LD, int SPT, int SG, int S>
__device__ __forceinline__ void ket_accumulate_all(const double (&R)[nherm(LA + LB + LC + LD)], const PairE<LC, LD>& Ecd,
                                                   double (&Hs)[SPT][nherm(LA + LB)]) {
    ket_accumulate<LA, LB, LC, LD, SG * SPT + S>(R, Ecd, Hs[S]);
    if constexpr (S + 1 < SPT) ket_accumulate_all<LA, LB, LC, LD, SPT, SG, S + 1>(R, Ecd, Hs);
}

// compile-time dispatch over the (warp-uniform) slab group
template <int LA, int LB, int LC, int LD, int NK, int SPT, int SG>
__device__ __forceinline__ void dispatch_ket(int sg, const double (&R)[nherm(LA + LB + LC + LD)], const PairE<LC, LD>& Ecd,
                                             double (&Hs)[SPT][nherm(LA + LB)]) {
    if (sg == SG) {
        ket_accumulate_all<LA, LB, LC, LD, SPT, SG, 0>(R, Ecd, Hs);
    } else if constexpr (SG + 1 < SlabCfg<LA, LB, LC, LD, NK, SPT>::G) {
        dispatch_ket<LA, LB, LC, LD, NK, SPT, SG + 1>(sg, R, Ecd, Hs);
    }
}
template <int LA, int LB, int LC, int LD, int NK, int SPT>
__global__ void __launch_bounds__(SlabCfg<LA, LB, LC, LD, NK, SPT>::BLOCK)
eri_jk_slab_kernel(PairGroup bra, PairGroup ket, BuildArgs a, int same_group) {
    using C = SlabCfg<LA, LB, LC, LD, NK, SPT>;
    using EL = E3Layout<LA, LB>;
    constexpr int NA = C::NA, NB = C::NB, NAB = C::NAB, L = C::L, NH = C::NH, G = C::G, NSUB = C::NSUB, BLOCK = C::BLOCK;
    extern __shared__ __align__(16) double smem[];
    __shared__ int off_s[NAB];
    const int ib_ = a.bra_list ? __ldg(a.bra_list + blockIdx.x) : (int)blockIdx.x;
    const double qab = __ldg(bra.Q + ib_);
    const int ket0 = blockIdx.y * a.ket_chunk;
    const int nket = ket_prefix_end(ket, a, qab, ib_, same_group, ket0);
    if (nket <= ket0) return;
    const AccMode fx{a.sc->fx_scale, a.red_eps};

    const int N = a.N, KAB = __ldg(bra.nprim + ib_);
    const int fa = __ldg(bra.fa + ib_), fb = __ldg(bra.fb + ib_);
    const int sa = __ldg(bra.sa + ib_), sb = __ldg(bra.sb + ib_);
    const double bra_deg = (sa == sb) ? 0.5 : 1.0;
    const float dab = __ldg(bra.Dp + ib_);

    // ---- shared-memory tables of the bra pair -------------------------------------------------------
    double* const tab = smem;                                  // [KAB][STRIDE]
    double* const pab_s = smem + (size_t)KAB * EL::STRIDE;     // [NAB] (+pad)
    double* const jab_all = pab_s + ((NAB + 1) & ~1);          // [NAB][BLOCK] when JSMEM
    {
        double ABx = 0, ABy = 0, ABz = 0;
        if constexpr (LB > 0) {
            ABx = __ldg(bra.AB + ib_); ABy = __ldg(bra.AB + bra.npair + ib_); ABz = __ldg(bra.AB + 2 * (size_t)bra.npair + ib_);
        }
        for (int t = threadIdx.x; t < 3 * KAB; t += BLOCK) {
            const int kb = t / 3, axis = t % 3;
            const double* base = bra.prim + (size_t)kb * PF_COUNT * bra.npair + ib_;
            const size_t np = bra.npair;
            const double p = __ldg(base + PF_P * np);
            const double xpa = __ldg(base + (PF_PAX + axis) * np);
            const double ab = axis == 0 ? ABx : (axis == 1 ? ABy : ABz);
            EAxis<LA, LB> E;
#pragma unroll
            for (int i = 0; i < EL::EA_N; ++i) E.e[i] = 0.0;
            E.build(0.5 / p, xpa, xpa + ab, false);
            double* dst = tab + (size_t)kb * EL::STRIDE;
#pragma unroll
            for (int i = 0; i < EL::EA_N; ++i) dst[EL::PRIM + axis * EL::EA_N + i] = E.e[i];
            if (axis == 0) {
                dst[0] = p; dst[1] = __ldg(base + PF_PX * np); dst[2] = __ldg(base + PF_PY * np);
                dst[3] = __ldg(base + PF_PZ * np); dst[4] = __ldg(base + PF_C * np);
            }
        }
        for (int i = threadIdx.x; i < NAB; i += BLOCK) pab_s[i] = __ldg(a.Pj + (size_t)(fa + i / NB) * N + fb + i % NB);
#pragma unroll
        for (int i = 0; i < NAB; ++i)
            if (threadIdx.x == (i % BLOCK)) off_s[i] = EL::off(i);     // compile-time constants
        __syncthreads();
        for (int idx = threadIdx.x; idx < KAB * NAB; idx += BLOCK) {
            const int kb = idx / NAB, iab = idx % NAB;
            const int ia = iab / NB, ibb = iab % NB;
            const int off = off_s[iab];
            double* tb = tab + (size_t)kb * EL::STRIDE;
            const double* ex = tb + EL::PRIM;
            const double* ey = ex + EL::EA_N;
            const double* ez = ey + EL::EA_N;
            double* dst = tb + EL::E3_OFF + off;
            const int ax = cart_x(LA, ia), ay = cart_y(LA, ia), az = cart_z(LA, ia);
            const int bx = cart_x(LB, ibb), by = cart_y(LB, ibb), bz = cart_z(LB, ibb);
            constexpr int NT = LA + LB + 1;
            int k = 0;
            for (int t = 0; t <= ax + bx; ++t)
                for (int u = 0; u <= ay + by; ++u)
                    for (int v = 0; v <= az + bz; ++v, ++k)
                        dst[k] = ex[(ax * (LB + 1) + bx) * NT + t] * ey[(ay * (LB + 1) + by) * NT + u] * ez[(az * (LB + 1) + bz) * NT + v];
            if (k & 1) dst[k] = 0.0;
        }
        if constexpr (C::JSMEM) {
            for (int i = threadIdx.x; i < NAB * BLOCK; i += BLOCK) jab_all[i] = 0.0;
        }
        __syncthreads();
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sg = blockIdx.z, ksub = warp;
    double* const jab_s = jab_all + threadIdx.x;

    double jab[NAB];
#pragma unroll
    for (int i = 0; i < NAB; ++i) jab[i] = 0.0;
    unsigned int nq = 0;

    // warp-level compaction of the surviving kets (see eri_jk_kernel): scan 32 candidates, queue the survivors,
    // run the quartet work on full warps
    constexpr int SW = 2;
    __shared__ int ket_queue[NSUB][32 * (SW + 1)];
    const float* const dsh_a = a.Dsh + (size_t)sa * a.nshell;     // rows of the shell-block maxima (global, L1-resident)
    const float* const dsh_b = a.Dsh + (size_t)sb * a.nshell;
    int* const queue = ket_queue[ksub];
    int qn = 0;
    const bool wide = a.ket_chunk >= 32 * NSUB * SW;      // see eri_jk_kernel: short chunks scan 32 candidates per step
    const int sw_eff = wide ? SW : 1;
    int scan = ket0 + ksub * (32 * sw_eff);
    while (true) {
        while (qn < 32 && scan < nket) {
            qn = wide ? scan_kets<SW>(ket, a.tau, scan, nket, qab, dab, dsh_a, dsh_b, queue, qn, lane)
                      : scan_kets<1>(ket, a.tau, scan, nket, qab, dab, dsh_a, dsh_b, queue, qn, lane);
            scan += 32 * NSUB * sw_eff;
        }
        __syncwarp();
        const int nrun = qn < 32 ? qn : 32;
        if (nrun == 0) break;
        if (lane < nrun) {
        const int ik_ = ket_queue[ksub][qn - nrun + lane];
        const int sc = __ldg(ket.sa + ik_), sd = __ldg(ket.sb + ik_);
        if (sg == 0) ++nq;
        const int fc = __ldg(ket.fa + ik_), fd = __ldg(ket.fb + ik_);
        double deg = bra_deg * ((sc == sd) ? 0.5 : 1.0);
        if (same_group && ik_ == ib_) deg *= 0.5;
        double CDx = 0, CDy = 0, CDz = 0;
        if constexpr (LD > 0) {
            CDx = __ldg(ket.AB + ik_); CDy = __ldg(ket.AB + ket.npair + ik_); CDz = __ldg(ket.AB + 2 * (size_t)ket.npair + ik_);
        }
        const int nkc = __ldg(ket.nprim + ik_);
        for (int kb = 0; kb < KAB; ++kb) {
            const double* __restrict__ tb = tab + (size_t)kb * EL::STRIDE;
            const double p = tb[0], Px = tb[1], Py = tb[2], Pz = tb[3], cP = tb[4] * deg;
            double Hs[SPT][NH];
#pragma unroll
            for (int s = 0; s < SPT; ++s)
#pragma unroll
                for (int i = 0; i < NH; ++i) Hs[s][i] = 0.0;
            for (int kc = 0; kc < nkc; ++kc) {
                double q, Qx, Qy, Qz, cQ;
                PairE<LC, LD> Ecd;
                load_prim<LC, LD>(ket, ik_, kc, CDx, CDy, CDz, q, Qx, Qy, Qz, cQ, Ecd, true);
                const double pq = p + q, rspq = fast_rsqrt(pq), rpq = rspq * rspq, alpha = p * q * rpq;
                const double X = Px - Qx, Y = Py - Qy, Z = Pz - Qz;
                double F[L + 1];
                boys<L>(alpha * (X * X + Y * Y + Z * Z), a.boys, F);
                double c[L + 1];
                {
                    double f = cP * cQ * rspq;
                    const double m2a = -2.0 * alpha;
#pragma unroll
                    for (int n = 0; n <= L; ++n) { c[n] = f * F[n]; f *= m2a; }
                }
                double R[nherm(L)];
                hermite_R<L, false>(c, X, Y, Z, R);
                dispatch_ket<LA, LB, LC, LD, NK, SPT, 0>(sg, R, Ecd, Hs);
            }
            slab_digest<LA, LB, LC, LD, NK, SPT>(sg, Hs, tb + EL::E3_OFF, pab_s, jab_s, jab, a, fx, fa, fb, fc, fd);
        }
        }   // lane < nrun
        qn -= nrun;
        __syncwarp();
    }

    // ---- J_ab: reduce over the CTA, one atomic per element ----
    __syncthreads();
    if constexpr (C::JSMEM) {
        for (int i = warp; i < NAB; i += BLOCK / 32) {
            double s = 0.0;
            for (int t = lane; t < BLOCK; t += 32) s += jab_all[i * BLOCK + t];
            s = warp_sum(s);
            if (lane == 0) red_add(a.AJ + (size_t)(fa + i / NB) * N + fb + i % NB, s, fx);
        }
    } else {
        double* red = smem;    // tables are dead now: [BLOCK/32][NAB]
#pragma unroll
        for (int i = 0; i < NAB; ++i) {
            const double s = warp_sum(jab[i]);
            if (lane == 0) red[warp * NAB + i] = s;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < NAB; i += BLOCK) {
            double s = 0.0;
            for (int w = 0; w < BLOCK / 32; ++w) s += red[w * NAB + i];
            red_add(a.AJ + (size_t)(fa + i / NB) * N + fb + i % NB, s, fx);
        }
    }
    nq = __reduce_add_sync(0xffffffffu, nq);
    if (lane == 0 && nq) atomicAdd(a.counter, (unsigned long long)nq);
}

template <int LA, int LB, int LC, int LD, int NK, int SPT>
inline size_t slab_smem_bytes(int KAB) {
    using C = SlabCfg<LA, LB, LC, LD, NK, SPT>;
    using EL = E3Layout<LA, LB>;
    size_t n = (size_t)KAB * EL::STRIDE + ((C::NAB + 1) & ~1) + (C::JSMEM ? (size_t)C::NAB * C::BLOCK : 0);
    const size_t red = (size_t)(C::BLOCK / 32) * C::NAB;
    if (n < red) n = red;
    return n * sizeof(double);
}

}  // namespace qcf
