// eri_class.cu -- one translation unit per angular class (LA LB | LC LD), compiled 21 times with
// -DQCF_LA= -DQCF_LB= -DQCF_LC= -DQCF_LD= so that the classes build in parallel.
#include "eri_device.cuh"

#ifndef QCF_LA
#error "compile with -DQCF_LA=.. -DQCF_LB=.. -DQCF_LC=.. -DQCF_LD=.."
#endif

namespace qcf {
namespace {
constexpr int LA = QCF_LA, LB = QCF_LB, LC = QCF_LC, LD = QCF_LD;

void launch_jk(int nk, int grid, int block, cudaStream_t s, const PairGroup& bra, const PairGroup& ket, const BuildArgs& a, int same) {
    if (nk == 1) eri_jk_kernel<LA, LB, LC, LD, 1><<<grid, block, 0, s>>>(bra, ket, a, same);
    else eri_jk_kernel<LA, LB, LC, LD, 2><<<grid, block, 0, s>>>(bra, ket, a, same);
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
