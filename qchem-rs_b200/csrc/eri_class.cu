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

// lanes per shell quartet the block kernel of this class is compiled for (highly contracted launches; the d
// shells of the supported basis sets are uncontracted or nearly so, so 8 is only built for the s/p bras)
constexpr int MAX_PS = USE_SLAB ? 1 : (LA <= 1 ? 8 : 4);

template <int NK>
void launch_slab(int nbra, int nket_max, int kpt, cudaStream_t s, const PairGroup& bra, const PairGroup& ket, BuildArgs a, int same) {
    if constexpr (USE_SLAB) {
        using C = SlabCfg<LA, LB, LC, LD, NK, SPT>;
        auto kern = eri_jk_slab_kernel<LA, LB, LC, LD, NK, SPT>;
        const size_t smem = slab_smem_bytes<LA, LB, LC, LD, NK, SPT>(bra.K);
        a.ket_chunk = 32 * C::NSUB * kpt;
        const dim3 grid(nbra, (nket_max + a.ket_chunk - 1) / a.ket_chunk, C::G);
        kern<<<grid, C::BLOCK, smem, s>>>(bra, ket, a, same);
    }
}

inline size_t block_smem_bytes(int braK, int nshell) {
    // staged bra primitives + two rows of the shell-block density maxima
    return (size_t)braK * BRA_S * sizeof(double) + 2 * (size_t)nshell * sizeof(float);
}

template <int NK, int PS>
void launch_block(int nbra, int nket_max, int block, int kpt, cudaStream_t s, const PairGroup& bra, const PairGroup& ket, BuildArgs a, int same) {
    if constexpr (!USE_SLAB && PS <= MAX_PS) {
        // candidates per CTA: kpt per thread, spread over PS lanes each; at least one 32-candidate scan step per warp
        a.ket_chunk = block * kpt / PS;
        if (a.ket_chunk < block) a.ket_chunk = block;
        const dim3 grid(nbra, (nket_max + a.ket_chunk - 1) / a.ket_chunk);
        const size_t smem = block_smem_bytes(bra.K, a.nshell);
        // the wide-scan instantiation only when every warp gets a full wide step out of the chunk
        if (BlockCfg<PS>::SW > 1 && a.ket_chunk >= block * BlockCfg<PS>::SW)
            eri_jk_kernel<LA, LB, LC, LD, NK, PS, true><<<grid, block, smem, s>>>(bra, ket, a, same);
        else
            eri_jk_kernel<LA, LB, LC, LD, NK, PS, false><<<grid, block, smem, s>>>(bra, ket, a, same);
    }
}

void launch_jk(int nk, int ps, int nbra, int nket_max, int block, int kpt, cudaStream_t s, const PairGroup& bra, const PairGroup& ket,
               const BuildArgs& a, int same) {
    if constexpr (USE_SLAB) {
        if (nk == 1) launch_slab<1>(nbra, nket_max, kpt, s, bra, ket, a, same);
        else launch_slab<2>(nbra, nket_max, kpt, s, bra, ket, a, same);
    } else {
        if (ps > MAX_PS) ps = MAX_PS;
        if (nk == 1) {
            if (ps >= 8) launch_block<1, 8>(nbra, nket_max, block, kpt, s, bra, ket, a, same);
            else if (ps >= 4) launch_block<1, 4>(nbra, nket_max, block, kpt, s, bra, ket, a, same);
            else launch_block<1, 1>(nbra, nket_max, block, kpt, s, bra, ket, a, same);
        } else {
            if (ps >= 8) launch_block<2, 8>(nbra, nket_max, block, kpt, s, bra, ket, a, same);
            else if (ps >= 4) launch_block<2, 4>(nbra, nket_max, block, kpt, s, bra, ket, a, same);
            else launch_block<2, 1>(nbra, nket_max, block, kpt, s, bra, ket, a, same);
        }
    }
}

// Opt in (once per device, at qcf_create) to the dynamic shared memory of the largest launch of this class.
template <class K>
cudaError_t set_smem(K kern, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
cudaError_t class_init(int max_bra_K, int nshell) {
    cudaError_t e = cudaSuccess;
    if constexpr (USE_SLAB) {
        e = set_smem(eri_jk_slab_kernel<LA, LB, LC, LD, 1, SPT>, slab_smem_bytes<LA, LB, LC, LD, 1, SPT>(max_bra_K));
        if (e == cudaSuccess) e = set_smem(eri_jk_slab_kernel<LA, LB, LC, LD, 2, SPT>, slab_smem_bytes<LA, LB, LC, LD, 2, SPT>(max_bra_K));
    } else {
        const size_t b = block_smem_bytes(max_bra_K, nshell);
        e = set_smem(eri_jk_kernel<LA, LB, LC, LD, 1, 1, true>, b);
        if (e == cudaSuccess) e = set_smem(eri_jk_kernel<LA, LB, LC, LD, 2, 1, true>, b);
        if (e == cudaSuccess) e = set_smem(eri_jk_kernel<LA, LB, LC, LD, 1, 1, false>, b);
        if (e == cudaSuccess) e = set_smem(eri_jk_kernel<LA, LB, LC, LD, 2, 1, false>, b);
        if constexpr (MAX_PS >= 4) {
            if (e == cudaSuccess) e = set_smem(eri_jk_kernel<LA, LB, LC, LD, 1, 4, false>, b);
            if (e == cudaSuccess) e = set_smem(eri_jk_kernel<LA, LB, LC, LD, 2, 4, false>, b);
        }
        if constexpr (MAX_PS >= 8) {
            if (e == cudaSuccess) e = set_smem(eri_jk_kernel<LA, LB, LC, LD, 1, 8, false>, b);
            if (e == cudaSuccess) e = set_smem(eri_jk_kernel<LA, LB, LC, LD, 2, 8, false>, b);
        }
    }
    return e;
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
    launch_jk, class_init, MAX_PS, launch_quartet,
#if QCF_LA == QCF_LC && QCF_LB == QCF_LD
    launch_schwarz
#else
    nullptr
#endif
};
}  // namespace qcf
