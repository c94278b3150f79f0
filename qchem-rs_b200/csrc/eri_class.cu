// eri_class.cu -- one translation unit per angular class (LA LB | LC LD), compiled 21 times with
// -DQCF_LA= -DQCF_LB= -DQCF_LC= -DQCF_LD= so that the classes build in parallel.
#include "eri_device.cuh"
#include "eri_slab.cuh"

#ifndef QCF_LA
#error "compile with -DQCF_LA=.. -DQCF_LB=.. -DQCF_LC=.. -DQCF_LD=.."
#endif

namespace qcf {
namespace {
constexpr int LA = QCF_LA, LB = QCF_LB, LC = QCF_LC, LD = QCF_LD;
constexpr int LTOT = LA + LB + LC + LD;
constexpr int NCD = ncart(LC) * ncart(LD);

// Which kernel serves this class, and how many ket components one thread of the slab kernel takes.
// Slab kernel: every class with a dp or dd bra whose R_tuv still fits the register file (L <= 6).
#ifdef QCF_SPT
constexpr int SPT = QCF_SPT;
#else
// largest divisor of NCD whose register-resident state (Hsum + R_tuv, or Hsum + digestion state) stays near 100 doubles
constexpr int choose_spt() {
    const int nh = nherm(LA + LB), nr = nherm(LTOT), nb = ncart(LB), nab = ncart(LA) * ncart(LB);
    int best = 1;
    for (int s = 1; s <= NCD; ++s) {
        if (NCD % s) continue;
        const int digest = s * 4 * (nb + 1) + (nab > 18 ? 0 : nab);
        const int live = s * nh + (nr > digest ? nr : digest);
        if (live <= 128) best = s;
    }
    return best;
}
constexpr int SPT = choose_spt();
#endif
#ifdef QCF_USE_SLAB
constexpr bool USE_SLAB = QCF_USE_SLAB;
#else
constexpr bool USE_SLAB = (LA == 2 && LB >= 1);   // every dp- and dd-bra class
#endif

template <int NK>
void launch_slab(int nbra, int nket_max, int kpt, cudaStream_t s, const PairGroup& bra, const PairGroup& ket, BuildArgs a, int same) {
    if constexpr (USE_SLAB) {
        using C = SlabCfg<LA, LB, LC, LD, NK, SPT>;
        auto kern = eri_jk_slab_kernel<LA, LB, LC, LD, NK, SPT>;
        const size_t smem = slab_smem_bytes<LA, LB, LC, LD, NK, SPT>(bra.K);
        // opt in to more than 48 KB of dynamic shared memory (per device and per function, so it is not cached in
        // a process-wide flag: one process may hold contexts on several devices)
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        a.ket_chunk = 32 * C::NSUB * kpt;
        const dim3 grid(nbra, (nket_max + a.ket_chunk - 1) / a.ket_chunk, C::G);
        kern<<<grid, C::BLOCK, smem, s>>>(bra, ket, a, same);
    }
}

void launch_jk(int nk, int nbra, int nket_max, int block, int kpt, cudaStream_t s, const PairGroup& bra, const PairGroup& ket,
               const BuildArgs& a0, int same) {
    BuildArgs a = a0;
    if constexpr (USE_SLAB) {
        if (nk == 1) launch_slab<1>(nbra, nket_max, kpt, s, bra, ket, a, same);
        else launch_slab<2>(nbra, nket_max, kpt, s, bra, ket, a, same);
    } else {
        a.ket_chunk = block * kpt;
        const dim3 grid(nbra, (nket_max + a.ket_chunk - 1) / a.ket_chunk);
        // dynamic shared memory: staged bra primitives + two rows of the shell-block density maxima
        const size_t smem = (size_t)bra.K * BRA_S * sizeof(double) + 2 * (size_t)a.nshell * sizeof(float);
        if (smem > 48 * 1024) {
            if (nk == 1) cudaFuncSetAttribute(eri_jk_kernel<LA, LB, LC, LD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            else cudaFuncSetAttribute(eri_jk_kernel<LA, LB, LC, LD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        if (nk == 1) eri_jk_kernel<LA, LB, LC, LD, 1><<<grid, block, smem, s>>>(bra, ket, a, same);
        else eri_jk_kernel<LA, LB, LC, LD, 2><<<grid, block, smem, s>>>(bra, ket, a, same);
    }
}
void launch_quartet(cudaStream_t s, const PairGroup& bra, int ib_, const PairGroup& ket, int ik_, const double* boys, double* out) {
    quartet_kernel<LA, LB, LC, LD><<<1, 32, 0, s>>>(bra, ib_, ket, ik_, boys, out);
}
#if QCF_LA == QCF_LC && QCF_LB == QCF_LD
void launch_schwarz(int grid, int block, cudaStream_t s, const PairGroup& g, const double* boys, double* Q) {
    schwarz_kernel<LA, LB><<<grid, block, 0, s>>>(g, boys, Q);
}
#endif
}  // namespace

#define QCF_CAT2(a, b, c, d) qcf_class_##a##b##c##d
#define QCF_CAT(a, b, c, d) QCF_CAT2(a, b, c, d)
extern "C" const ClassLaunch QCF_CAT(QCF_LA, QCF_LB, QCF_LC, QCF_LD) = {
    launch_jk, launch_quartet,
#if QCF_LA == QCF_LC && QCF_LB == QCF_LD
    launch_schwarz
#else
    nullptr
#endif
};
}  // namespace qcf
