// onee.cuh -- one-electron integrals (overlap, kinetic, nuclear attraction) on the device.
//
// Stands in for molint::overlap / molint::kinetic / molint::nuclear (call sites core/src/hf/rhf.rs:41-43,
// uhf.rs:52-54; the crate is absent from the reference tree).  SURVEY.md 8f item 1: with these three the
// Rust driver no longer needs molint at all.  Same Hermite machinery as the two-electron code:
//   S_ab = c (pi/p)^(3/2) E0x E0y E0z
//   T_ab = -2b^2 S(j+2) + b(2j+1) S(j) - j(j-1)/2 S(j-2)   per axis
//   V_ab = -sum_C Z_C (2 pi/p) c sum_tuv E_t E_u E_v R_tuv(p, P - C)
// One warp per shell pair; lanes stride over (primitive pair, nucleus) items; warp-shuffle reduction.
#pragma once
#include "eri_device.cuh"

namespace qcf {

struct ShellData {
    int nshell, natoms, N;
    const int* atom;      // [nshell]
    const int* l;         // [nshell]
    const int* nprim;     // [nshell]
    const int* prim_off;  // [nshell]
    const int* off;       // [nshell+1] first basis function
    const double* exps;
    const double* coefs;
    const double* xyz;    // [3*natoms]
    const double* charge; // [natoms]
    const double* fscale; // [N] per-component normalisation factor
};

template <int LA, int LB>
__device__ void onee_pair(const ShellData& sd, int sa, int sb, const double* __restrict__ boys_table,
                          double* __restrict__ S, double* __restrict__ T, double* __restrict__ V) {
    constexpr int NA = ncart(LA), NB = ncart(LB), NAB = NA * NB, L = LA + LB;
    const int lane = threadIdx.x & 31;
    const double* A = sd.xyz + 3 * sd.atom[sa];
    const double* B = sd.xyz + 3 * sd.atom[sb];
    const double ABx = A[0] - B[0], ABy = A[1] - B[1], ABz = A[2] - B[2];
    const double AB2 = ABx * ABx + ABy * ABy + ABz * ABz;
    const int npa = sd.nprim[sa], npb = sd.nprim[sb];
    const int nitem = npa * npb * sd.natoms;
    double accS[NAB], accT[NAB], accV[NAB];
#pragma unroll
    for (int i = 0; i < NAB; ++i) { accS[i] = 0.0; accT[i] = 0.0; accV[i] = 0.0; }

    for (int item = lane; item < nitem; item += 32) {
        const int kp = item / sd.natoms, ic = item % sd.natoms;
        const int ka = kp / npb, kb = kp % npb;
        const double a = sd.exps[sd.prim_off[sa] + ka], b = sd.exps[sd.prim_off[sb] + kb];
        const double p = a + b, rp = 1.0 / p, mu = a * b * rp;
        const double c = sd.coefs[sd.prim_off[sa] + ka] * sd.coefs[sd.prim_off[sb] + kb] * exp(-mu * AB2);
        const double Px = (a * A[0] + b * B[0]) * rp, Py = (a * A[1] + b * B[1]) * rp, Pz = (a * A[2] + b * B[2]) * rp;
        EAxis<LA, LB + 2> E[3];
        E[0].build(0.5 * rp, Px - A[0], Px - B[0], false);
        E[1].build(0.5 * rp, Py - A[1], Py - B[1], false);
        E[2].build(0.5 * rp, Pz - A[2], Pz - B[2], false);
        if (ic == 0) {
            const double s3 = c * (PI_D * rp) * sqrt(PI_D * rp);
#pragma unroll
            for (int ia = 0; ia < NA; ++ia)
#pragma unroll
                for (int ib = 0; ib < NB; ++ib) {
                    double s1[3], t1[3];
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        const int i = cart_pow(LA, ia, ax), j = cart_pow(LB, ib, ax);
                        s1[ax] = E[ax].get(i, j, 0);
                        double t = -2.0 * b * b * E[ax].get(i, j + 2, 0) + b * (2 * j + 1) * s1[ax];
                        if (j >= 2) t -= 0.5 * j * (j - 1) * E[ax].get(i, j - 2, 0);
                        t1[ax] = t;
                    }
                    accS[ia * NB + ib] = fma(s3, s1[0] * s1[1] * s1[2], accS[ia * NB + ib]);
                    accT[ia * NB + ib] = fma(s3, t1[0] * s1[1] * s1[2] + s1[0] * t1[1] * s1[2] + s1[0] * s1[1] * t1[2],
                                             accT[ia * NB + ib]);
                }
        }
        {
            const double* C = sd.xyz + 3 * ic;
            const double X = Px - C[0], Y = Py - C[1], Z = Pz - C[2];
            double F[L + 1];
            boys<L>(p * (X * X + Y * Y + Z * Z), boys_table, F);
            double cn[L + 1];
            double f = -sd.charge[ic] * 2.0 * PI_D * rp * c;
#pragma unroll
            for (int n = 0; n <= L; ++n) { cn[n] = f * F[n]; f *= -2.0 * p; }
            double R[nherm(L)];
            hermite_R<L, false>(cn, X, Y, Z, R);
#pragma unroll
            for (int ia = 0; ia < NA; ++ia)
#pragma unroll
                for (int ib = 0; ib < NB; ++ib) {
                    const int ax = cart_x(LA, ia), ay = cart_y(LA, ia), az = cart_z(LA, ia);
                    const int bx = cart_x(LB, ib), by = cart_y(LB, ib), bz = cart_z(LB, ib);
                    double s = 0.0;
#pragma unroll
                    for (int t = 0; t <= ax + bx; ++t)
#pragma unroll
                        for (int u = 0; u <= ay + by; ++u) {
                            const double exy = E[0].get(ax, bx, t) * E[1].get(ay, by, u);
#pragma unroll
                            for (int v = 0; v <= az + bz; ++v) s = fma(exy * E[2].get(az, bz, v), R[hidx(t, u, v)], s);
                        }
                    accV[ia * NB + ib] += s;
                }
        }
    }
    const int fa = sd.off[sa], fb = sd.off[sb], N = sd.N;
#pragma unroll
    for (int i = 0; i < NAB; ++i) {
        const double s = warp_sum(accS[i]), t = warp_sum(accT[i]), v = warp_sum(accV[i]);
        if (lane == 0) {
            const int ia = fa + i / NB, ib = fb + i % NB;
            const double sc = sd.fscale[ia] * sd.fscale[ib];
            S[(size_t)ia * N + ib] = sc * s; S[(size_t)ib * N + ia] = sc * s;
            T[(size_t)ia * N + ib] = sc * t; T[(size_t)ib * N + ia] = sc * t;
            V[(size_t)ia * N + ib] = sc * v; V[(size_t)ib * N + ia] = sc * v;
        }
    }
}

// one warp per shell pair (s1 >= s2); the higher angular momentum is put first
__global__ void __launch_bounds__(128) onee_kernel(ShellData sd, const double* __restrict__ boys_table, double* __restrict__ S,
                                                   double* __restrict__ T, double* __restrict__ V) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long npair = (long long)sd.nshell * (sd.nshell + 1) / 2;
    if (w >= npair) return;
    int s1 = (int)((sqrt(8.0 * (double)w + 1.0) - 1.0) * 0.5);
    while ((long long)(s1 + 1) * (s1 + 2) / 2 <= w) ++s1;
    while ((long long)s1 * (s1 + 1) / 2 > w) --s1;
    int s2 = (int)(w - (long long)s1 * (s1 + 1) / 2);
    if (sd.l[s2] > sd.l[s1]) { const int t = s1; s1 = s2; s2 = t; }
    switch (sd.l[s1] * 3 + sd.l[s2]) {
        case 0: onee_pair<0, 0>(sd, s1, s2, boys_table, S, T, V); break;
        case 3: onee_pair<1, 0>(sd, s1, s2, boys_table, S, T, V); break;
        case 4: onee_pair<1, 1>(sd, s1, s2, boys_table, S, T, V); break;
        case 6: onee_pair<2, 0>(sd, s1, s2, boys_table, S, T, V); break;
        case 7: onee_pair<2, 1>(sd, s1, s2, boys_table, S, T, V); break;
        case 8: onee_pair<2, 2>(sd, s1, s2, boys_table, S, T, V); break;
    }
}

}  // namespace qcf
