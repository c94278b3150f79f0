// engine.cu -- host side of the Fock-build engine and the C ABI of include/qcfock.h.
//
// qcf_create   : shell pairs -> (la,lb,K) groups -> Schwarz factors on the GPU -> sorted SoA pair data in HBM
//                (replaces the one-off molint::eri call, core/src/hf/rhf.rs:45, uhf.rs:55)
// qcf_build_*  : density scaling + shell-block maxima -> one eri_jk launch per (bra group, ket group) ->
//                symmetrise / combine  (replaces rhf.rs:58-62,152-167 and uhf.rs:210-227)
// There is no CPU fallback anywhere in this file: every failure of the CUDA runtime is reported.
#include "../../include/qcfock.h"
#include "eri_device.cuh"
#include "onee.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <string>
#include <vector>

using namespace qcf;

// the 21 angular classes (bra pair class >= ket pair class), one object file each
#define QCF_DECL(a, b, c, d) extern "C" const ClassLaunch qcf_class_##a##b##c##d;
#define QCF_ALL_CLASSES(X) \
    X(0, 0, 0, 0) \
    X(1, 0, 0, 0) X(1, 0, 1, 0) \
    X(1, 1, 0, 0) X(1, 1, 1, 0) X(1, 1, 1, 1) \
    X(2, 0, 0, 0) X(2, 0, 1, 0) X(2, 0, 1, 1) X(2, 0, 2, 0) \
    X(2, 1, 0, 0) X(2, 1, 1, 0) X(2, 1, 1, 1) X(2, 1, 2, 0) X(2, 1, 2, 1) \
    X(2, 2, 0, 0) X(2, 2, 1, 0) X(2, 2, 1, 1) X(2, 2, 2, 0) X(2, 2, 2, 1) X(2, 2, 2, 2)
QCF_ALL_CLASSES(QCF_DECL)

namespace {

constexpr int NPAIRCLASS = 6;
inline int pair_class(int la, int lb) { return la * (la + 1) / 2 + lb; }  // ss0 ps1 pp2 ds3 dp4 dd5

const ClassLaunch* class_table(int bra_cls, int ket_cls) {
    static const ClassLaunch* tab[NPAIRCLASS][NPAIRCLASS] = {};
    static bool init = false;
    if (!init) {
#define QCF_REG(a, b, c, d) tab[pair_class(a, b)][pair_class(c, d)] = &qcf_class_##a##b##c##d;
        QCF_ALL_CLASSES(QCF_REG)
        init = true;
    }
    return tab[bra_cls][ket_cls];
}

// SURVEY.md 8d op-count model per primitive quartet of class (la lb | lc ld)
double model_flops_prim(int la, int lb, int lc, int ld) {
    const int Lab = la + lb, Lcd = lc + ld, L = Lab + Lcd;
    const int nab = ncart(la) * ncart(lb), ncd = ncart(lc) * ncart(ld);
    double f = (20 + 3 * L) + 30;
    for (int n = 0; n < L; ++n) f += 3.0 * nherm(L - n);
    f += 2.0 * nherm(Lab) * nherm(Lcd) * nab;
    f += 2.0 * nherm(Lcd) * nab * ncd;
    return f;
}

struct HostPair {
    int sa, sb;       // shell ids, shell sa has l >= shell sb
    double Q;
    int keff = -1;                      // primitive pairs kept (-1: all K)
    std::vector<unsigned char> order;   // primitive-pair order (by primitive Schwarz factor); empty: natural
};

struct Group {
    int la, lb, K, cls;
    std::vector<HostPair> pairs;
    // device
    int* d_fa = nullptr; int* d_fb = nullptr; int* d_sa = nullptr; int* d_sb = nullptr; int* d_np = nullptr;
    double* d_Q = nullptr; double* d_Qb = nullptr; double* d_prim = nullptr; double* d_AB = nullptr;
    PairGroup dev{};
};

}  // namespace

struct qcf_ctx {
    std::string err;
    int device = 0, rank = 0, world = 1, block = 64, kets_per_thread = 32, target_ctas = 296;
    double serial_cap = 4e6;          // model flops one thread may run serially in one launch
    double tau = 1e-12;
    bool screening = true;
    // basis (host copies)
    int natoms = 0, nshell = 0, N = 0;
    std::vector<double> xyz, exps, coefs, charge;
    std::vector<int> sh_atom, sh_l, sh_np, sh_po, sh_off;
    std::vector<double> fscale;       // per basis function component scale
    std::vector<Group> groups;
    std::map<std::pair<int, int>, std::pair<int, int>> pair_index;  // (sa,sb) canonical -> (group, index)
    double qmax = 0;
    long long prim_total = 0, prim_kept = 0;
    // device state
    double* d_boys = nullptr;
    double* d_fscale = nullptr;
    int* d_shoff = nullptr;
    double *d_Pin[2] = {nullptr, nullptr}, *d_Pj = nullptr, *d_Pk[2] = {nullptr, nullptr};
    double *d_AJ = nullptr, *d_AK[2] = {nullptr, nullptr}, *d_G[2] = {nullptr, nullptr};
    float* d_Dsh = nullptr;
    unsigned int* d_dmax = nullptr;
    unsigned long long* d_counters = nullptr;
    int max_launch = 0;
    double* h_pin = nullptr;  // pinned staging, 2*N*N
    static constexpr int MAXSTREAM = 64;
    int nstreams = 8;
    cudaStream_t streams[MAXSTREAM] = {};
    cudaStream_t main_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join[MAXSTREAM] = {}, ev_t0 = nullptr, ev_t1 = nullptr, ev_h0 = nullptr, ev_h1 = nullptr;
    // stats of the last build
    struct LaunchRec { int bra, ket; float ms = 0; };
    bool profile = false;                 // QCF_PROFILE=1: serialise the class launches and time each one
    std::vector<cudaEvent_t> prof_ev;
    std::vector<LaunchRec> launches;
    std::vector<unsigned long long> launch_cnt;
    qcf_stats_t stats{};
    bool counters_pending = false;
};

namespace {

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                         \
            return QCF_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

// ---- Boys table (host, long double) ------------------------------------------------------------------
long double boys_series(int m, long double T) {
    long double term = 1.0L / (2 * m + 1), sum = term;
    for (int k = 1; k < 600; ++k) {
        term *= 2.0L * T / (2 * m + 2 * k + 1);
        sum += term;
        if (term < 1e-24L * sum) break;
    }
    return expl(-T) * sum;
}

std::vector<double> make_boys_table() {
    const int MR = BOYS_LTOT + BOYS_ORDER + 1;
    std::vector<double> tab((size_t)(BOYS_LTOT + 1) * BOYS_NGRID * BOYS_ROW, 0.0);
    std::vector<long double> F(MR);
    for (int g = 0; g < BOYS_NGRID; ++g) {
        const long double T0 = (long double)g / BOYS_PER_UNIT;
        F[MR - 1] = boys_series(MR - 1, T0);
        const long double e = expl(-T0);
        for (int m = MR - 1; m > 0; --m) F[m - 1] = (2.0L * T0 * F[m] + e) / (2 * m - 1);
        for (int L = 0; L <= BOYS_LTOT; ++L) {
            long double fact = 1.0L;
            for (int k = 0; k <= BOYS_ORDER; ++k) {
                if (k > 0) fact *= k;
                tab[((size_t)L * BOYS_NGRID + g) * BOYS_ROW + k] = (double)(F[L + k] / fact);
            }
            tab[((size_t)L * BOYS_NGRID + g) * BOYS_ROW + BOYS_ORDER + 1] = (double)e;
        }
    }
    return tab;
}

// ---- small kernels --------------------------------------------------------------------------------
__global__ void scale_density_kernel(int N, const double* __restrict__ fs, const double* __restrict__ Pa,
                                     const double* __restrict__ Pb, double* __restrict__ Pk0, double* __restrict__ Pk1,
                                     double* __restrict__ Pj) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * N) return;
    const int i = idx / N, j = idx % N;
    const double s = fs[i] * fs[j];
    const double a = Pa[idx] * s;
    Pk0[idx] = a;
    if (Pb) {
        const double b = Pb[idx] * s;
        Pk1[idx] = b;
        Pj[idx] = a + b;
    }
}

// one thread per shell block: max |P| over the block and over all densities
__global__ void dens_block_max_kernel(int N, int nshell, const int* __restrict__ shoff, const double* __restrict__ P0,
                                      const double* __restrict__ P1, const double* __restrict__ P2, float* __restrict__ Dsh,
                                      unsigned int* __restrict__ dmax) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (idx < nshell * nshell) {
        const int sa = idx / nshell, sb = idx % nshell;
        double mm = 0.0;
        for (int i = shoff[sa]; i < shoff[sa + 1]; ++i)
            for (int j = shoff[sb]; j < shoff[sb + 1]; ++j) {
                const size_t k = (size_t)i * N + j;
                mm = fmax(mm, fabs(P0[k]));
                if (P1) mm = fmax(mm, fabs(P1[k]));
                if (P2) mm = fmax(mm, fabs(P2[k]));
            }
        // round up so that the float bound never undercuts the double value
        m = __double2float_ru(mm);
        Dsh[idx] = m;
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(dmax, __float_as_uint(m));
}

// mode 0: G0 = fs fs (2(AJ+AJ^T) - 1/2 (AK0+AK0^T))                       (RHF)
// mode 1: G0/G1 = fs fs (2(AJ+AJ^T) - (AKs + AKs^T))                      (UHF)
// mode 2: G0 = fs fs 2(AJ+AJ^T),  G1 = fs fs (AK0+AK0^T)                  (J and K)
__global__ void finalize_kernel(int N, int mode, const double* __restrict__ fs, const double* __restrict__ AJ,
                                const double* __restrict__ AK0, const double* __restrict__ AK1, double* __restrict__ G0,
                                double* __restrict__ G1) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * N) return;
    const int i = idx / N, j = idx % N;
    const size_t tr = (size_t)j * N + i;
    const double s = fs[i] * fs[j];
    const double J = 2.0 * (AJ[idx] + AJ[tr]);
    const double K0 = AK0[idx] + AK0[tr];
    if (mode == 0) {
        G0[idx] = s * (J - 0.5 * K0);
    } else if (mode == 1) {
        const double K1 = AK1[idx] + AK1[tr];
        G0[idx] = s * (J - K0);
        G1[idx] = s * (J - K1);
    } else {
        G0[idx] = s * J;
        G1[idx] = s * K0;
    }
}

template <int L>
__global__ void boys_test_kernel(int n, const double* __restrict__ T, const double* __restrict__ table, double* __restrict__ F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double f[L + 1];
    boys<L>(T[i], table, f);
#pragma unroll
    for (int m = 0; m <= L; ++m) F[(size_t)i * (L + 1) + m] = f[m];
}

// FP64 FMA peak: 8 independent dependent chains per thread
__global__ void fp64_peak_kernel(double* out, int iters, double x) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double y = x * 0.5;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

// ---- pair construction -------------------------------------------------------------------------------
struct PairArrays {
    std::vector<int> fa, fb, sa, sb, np;
    std::vector<double> Q, Qb, prim, AB;
};

// power-of-two ceiling of a Schwarz factor: Q <= 2^ceil(log2 Q)
inline int q_bucket(double Q) { return Q > 0 ? (int)std::ceil(std::log2(Q)) : -100000; }
inline double q_bucket_ceiling(double Q) { return Q > 0 ? std::ldexp(1.0, q_bucket(Q)) : 0.0; }

void fill_pair_arrays(const qcf_ctx* c, const Group& g, PairArrays& out) {
    const size_t np = g.pairs.size();
    out.fa.resize(np); out.fb.resize(np); out.sa.resize(np); out.sb.resize(np); out.Q.resize(np); out.Qb.resize(np); out.np.resize(np);
    out.prim.assign((size_t)g.K * PF_COUNT * np, 0.0);
    out.AB.assign(3 * np, 0.0);
    const double cpi = std::sqrt(2.0) * std::pow(PI_D, 1.25);
    for (size_t i = 0; i < np; ++i) {
        const int sa = g.pairs[i].sa, sb = g.pairs[i].sb;
        out.fa[i] = c->sh_off[sa]; out.fb[i] = c->sh_off[sb]; out.sa[i] = sa; out.sb[i] = sb; out.Q[i] = g.pairs[i].Q; out.Qb[i] = q_bucket_ceiling(g.pairs[i].Q);
        const double* A = &c->xyz[3 * c->sh_atom[sa]];
        const double* B = &c->xyz[3 * c->sh_atom[sb]];
        double AB2 = 0;
        for (int k = 0; k < 3; ++k) { out.AB[k * np + i] = A[k] - B[k]; AB2 += (A[k] - B[k]) * (A[k] - B[k]); }
        out.np[i] = g.pairs[i].keff < 0 ? g.K : g.pairs[i].keff;
        const int npb = c->sh_np[sb];
        for (int kk = 0; kk < g.K; ++kk) {
            {
                const int ksel = g.pairs[i].order.empty() ? kk : g.pairs[i].order[kk];
                const int ia = ksel / npb, ib = ksel % npb;
                const double a = c->exps[c->sh_po[sa] + ia], b = c->exps[c->sh_po[sb] + ib];
                const double ca = c->coefs[c->sh_po[sa] + ia], cb = c->coefs[c->sh_po[sb] + ib];
                const double p = a + b, mu = a * b / p;
                double* f = &out.prim[(size_t)kk * PF_COUNT * np + i];
                f[PF_P * np] = p;
                for (int k = 0; k < 3; ++k) {
                    const double Pk = (a * A[k] + b * B[k]) / p;
                    f[(PF_PX + k) * np] = Pk;
                    f[(PF_PAX + k) * np] = Pk - A[k];
                }
                f[PF_C * np] = cpi * ca * cb * std::exp(-mu * AB2) / p;
            }
        }
    }
}

template <class T>
cudaError_t upload(T** dptr, const std::vector<T>& h) {
    cudaError_t e = cudaMalloc((void**)dptr, std::max<size_t>(h.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (h.empty()) return cudaSuccess;
    return cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

void free_group(Group& g) {
    cudaFree(g.d_fa); cudaFree(g.d_fb); cudaFree(g.d_sa); cudaFree(g.d_sb); cudaFree(g.d_np); cudaFree(g.d_Q); cudaFree(g.d_Qb); cudaFree(g.d_prim); cudaFree(g.d_AB);
    g.d_fa = g.d_fb = g.d_sa = g.d_sb = g.d_np = nullptr; g.d_Q = g.d_Qb = g.d_prim = g.d_AB = nullptr;
}

int upload_group(qcf_ctx* ctx, Group& g) {
    PairArrays pa;
    fill_pair_arrays(ctx, g, pa);
    free_group(g);
    CK(upload(&g.d_fa, pa.fa)); CK(upload(&g.d_fb, pa.fb)); CK(upload(&g.d_sa, pa.sa)); CK(upload(&g.d_sb, pa.sb)); CK(upload(&g.d_np, pa.np));
    CK(upload(&g.d_Q, pa.Q)); CK(upload(&g.d_Qb, pa.Qb)); CK(upload(&g.d_prim, pa.prim)); CK(upload(&g.d_AB, pa.AB));
    g.dev.npair = (int)g.pairs.size(); g.dev.K = g.K; g.dev.la = g.la; g.dev.lb = g.lb;
    g.dev.fa = g.d_fa; g.dev.fb = g.d_fb; g.dev.sa = g.d_sa; g.dev.sb = g.d_sb; g.dev.nprim = g.d_np; g.dev.Q = g.d_Q; g.dev.Qb = g.d_Qb; g.dev.prim = g.d_prim; g.dev.AB = g.d_AB;
    return QCF_OK;
}

int build_pairs(qcf_ctx* ctx) {
    const int ns = ctx->nshell;
    std::map<std::tuple<int, int, int>, int> gid;   // (cls, K) -> group
    // distance / overlap prescreen (only when screening is on)
    for (int s1 = 0; s1 < ns; ++s1)
        for (int s2 = 0; s2 <= s1; ++s2) {
            int sa = s1, sb = s2;
            if (ctx->sh_l[sb] > ctx->sh_l[sa]) std::swap(sa, sb);
            const int la = ctx->sh_l[sa], lb = ctx->sh_l[sb];
            if (ctx->screening) {
                const double* A = &ctx->xyz[3 * ctx->sh_atom[sa]];
                const double* B = &ctx->xyz[3 * ctx->sh_atom[sb]];
                double R2 = 0;
                for (int k = 0; k < 3; ++k) R2 += (A[k] - B[k]) * (A[k] - B[k]);
                double est = 0;
                for (int ia = 0; ia < ctx->sh_np[sa]; ++ia)
                    for (int ib = 0; ib < ctx->sh_np[sb]; ++ib) {
                        const double a = ctx->exps[ctx->sh_po[sa] + ia], b = ctx->exps[ctx->sh_po[sb] + ib];
                        const double p = a + b;
                        est = std::max(est, std::fabs(ctx->coefs[ctx->sh_po[sa] + ia] * ctx->coefs[ctx->sh_po[sb] + ib]) *
                                                std::pow(PI_D / p, 1.5) * std::exp(-a * b / p * R2));
                    }
                est *= std::pow(1.0 + std::sqrt(R2), la + lb);
                if (est < 1e-18) continue;
            }
            const int K = ctx->sh_np[sa] * ctx->sh_np[sb];
            const int cls = pair_class(la, lb);
            auto key = std::make_tuple(cls, K, 0);
            auto it = gid.find(key);
            if (it == gid.end()) {
                Group g; g.la = la; g.lb = lb; g.K = K; g.cls = cls;
                ctx->groups.push_back(std::move(g));
                it = gid.emplace(key, (int)ctx->groups.size() - 1).first;
            }
            ctx->groups[it->second].pairs.push_back({sa, sb, 0.0});
        }
    std::sort(ctx->groups.begin(), ctx->groups.end(), [](const Group& x, const Group& y) {
        return x.cls != y.cls ? x.cls < y.cls : x.K < y.K;
    });
    // Schwarz factors on the device, group by group
    ctx->qmax = 0;
    for (auto& g : ctx->groups) {
        int rc = upload_group(ctx, g);
        if (rc) return rc;
        double* dQ = nullptr;
        const int np = (int)g.pairs.size();
        CK(cudaMalloc(&dQ, sizeof(double) * np));
        const ClassLaunch* cl = class_table(g.cls, g.cls);
        cl->schwarz((np + 63) / 64, 64, 0, g.dev, ctx->d_boys, dQ);
        CK(cudaGetLastError());
        std::vector<double> Q(np);
        CK(cudaMemcpy(Q.data(), dQ, sizeof(double) * np, cudaMemcpyDeviceToHost));
        cudaFree(dQ);
        for (int i = 0; i < np; ++i) { g.pairs[i].Q = Q[i]; ctx->qmax = std::max(ctx->qmax, Q[i]); }
    }
    // primitive screening: Schwarz factor of every primitive pair on its own (the same class kernel on a
    // K = 1 list); inside each shell pair the primitives are sorted by it and those that cannot contribute
    // Q_k * Q_max >= 1e-4 tau to any integral are dropped (the kernels loop over nprim[i] <= K primitives)
    ctx->prim_total = ctx->prim_kept = 0;
    for (auto& g : ctx->groups) {
        const long long np = (long long)g.pairs.size();
        ctx->prim_total += np * g.K;
        if (!ctx->screening || g.K == 1) { ctx->prim_kept += np * g.K; continue; }
        Group g1; g1.la = g.la; g1.lb = g.lb; g1.K = 1; g1.cls = g.cls;
        g1.pairs.reserve((size_t)np * g.K);
        for (const auto& pr : g.pairs)
            for (int k = 0; k < g.K; ++k) { HostPair h{pr.sa, pr.sb, 0.0}; h.order.assign(1, (unsigned char)k); g1.pairs.push_back(std::move(h)); }
        // the K = 1 list holds primitive `order[0]` of each pair in slot 0
        int rc = upload_group(ctx, g1);
        if (rc) return rc;
        const int n1 = (int)g1.pairs.size();
        double* dQ = nullptr;
        CK(cudaMalloc(&dQ, sizeof(double) * n1));
        class_table(g.cls, g.cls)->schwarz((n1 + 63) / 64, 64, 0, g1.dev, ctx->d_boys, dQ);
        CK(cudaGetLastError());
        std::vector<double> Qk(n1);
        CK(cudaMemcpy(Qk.data(), dQ, sizeof(double) * n1, cudaMemcpyDeviceToHost));
        cudaFree(dQ);
        free_group(g1);
        const double pcut = ctx->tau * 1e-4 / std::max(ctx->qmax, 1e-300);
        for (long long i = 0; i < np; ++i) {
            auto& pr = g.pairs[i];
            pr.order.resize(g.K);
            for (int k = 0; k < g.K; ++k) pr.order[k] = (unsigned char)k;
            const double* q = &Qk[(size_t)i * g.K];
            std::stable_sort(pr.order.begin(), pr.order.end(), [&](unsigned char x, unsigned char y) { return q[x] > q[y]; });
            int keep = 0;
            while (keep < g.K && q[pr.order[keep]] >= pcut) ++keep;
            pr.keff = std::max(keep, 1);
            ctx->prim_kept += pr.keff;
        }
    }
    // drop negligible pairs, sort by Q, final upload
    size_t npairs = 0;
    for (auto& g : ctx->groups) {
        if (ctx->screening) {
            const double cut = ctx->tau * 1e-2 / std::max(ctx->qmax, 1e-300);
            g.pairs.erase(std::remove_if(g.pairs.begin(), g.pairs.end(), [&](const HostPair& p) { return p.Q < cut; }), g.pairs.end());
        }
        std::stable_sort(g.pairs.begin(), g.pairs.end(), [](const HostPair& x, const HostPair& y) {
            const int bx = q_bucket(x.Q), by = q_bucket(y.Q);
            if (bx != by) return bx > by;
            return x.sa != y.sa ? x.sa < y.sa : x.sb < y.sb;
        });
        npairs += g.pairs.size();
    }
    ctx->groups.erase(std::remove_if(ctx->groups.begin(), ctx->groups.end(), [](const Group& g) { return g.pairs.empty(); }),
                      ctx->groups.end());
    for (size_t gi = 0; gi < ctx->groups.size(); ++gi) {
        auto& g = ctx->groups[gi];
        int rc = upload_group(ctx, g);
        if (rc) return rc;
        for (size_t i = 0; i < g.pairs.size(); ++i) ctx->pair_index[{g.pairs[i].sa, g.pairs[i].sb}] = {(int)gi, (int)i};
    }
    ctx->stats.n_pairs = (int)npairs;
    ctx->stats.n_groups = (int)ctx->groups.size();
    ctx->stats.prim_pairs = ctx->prim_total;
    ctx->stats.prim_pairs_kept = ctx->prim_kept;
    return QCF_OK;
}

// ---- the build ---------------------------------------------------------------------------------------
// dPa/dPb: unscaled densities on the device; results into dG0/dG1 (device).  mode as finalize_kernel.
int run_build(qcf_ctx* ctx, int mode, const double* dPa, const double* dPb, double* dG0, double* dG1, cudaStream_t user) {
    const int N = ctx->N, ns = ctx->nshell;
    const size_t nn = (size_t)N * N;
    const int nk = (mode == 1) ? 2 : 1;
    cudaStream_t ms = user;
    const int tpb = 256;
    const int nblk = (int)((nn + tpb - 1) / tpb);
    CK(cudaEventRecord(ctx->ev_t0, ms));
    scale_density_kernel<<<nblk, tpb, 0, ms>>>(N, ctx->d_fscale, dPa, nk == 2 ? dPb : nullptr, ctx->d_Pk[0], ctx->d_Pk[1], ctx->d_Pj);
    const double* Pj = nk == 2 ? ctx->d_Pj : ctx->d_Pk[0];
    CK(cudaMemsetAsync(ctx->d_dmax, 0, sizeof(unsigned int), ms));
    dens_block_max_kernel<<<(ns * ns + 255) / 256, 256, 0, ms>>>(N, ns, ctx->d_shoff, ctx->d_Pk[0], nk == 2 ? ctx->d_Pk[1] : nullptr,
                                                                 nk == 2 ? ctx->d_Pj : nullptr, ctx->d_Dsh, ctx->d_dmax);
    CK(cudaMemsetAsync(ctx->d_AJ, 0, nn * sizeof(double), ms));
    CK(cudaMemsetAsync(ctx->d_AK[0], 0, nn * sizeof(double), ms));
    if (nk == 2) CK(cudaMemsetAsync(ctx->d_AK[1], 0, nn * sizeof(double), ms));
    CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long) * ctx->max_launch, ms));
    // the global density maximum steers the ket cutoffs; it is needed on the host for nothing else, so
    // read it back once (tiny, synchronous on this stream)
    unsigned int dmax_bits = 0;
    CK(cudaMemcpyAsync(&dmax_bits, ctx->d_dmax, sizeof(unsigned int), cudaMemcpyDeviceToHost, ms));
    CK(cudaStreamSynchronize(ms));
    float dmaxf;
    std::memcpy(&dmaxf, &dmax_bits, sizeof(float));
    BuildArgs a{};
    a.N = N; a.nshell = ns; a.nk = nk;
    a.Pj = Pj; a.Pk0 = ctx->d_Pk[0]; a.Pk1 = ctx->d_Pk[1];
    a.AJ = ctx->d_AJ; a.AK0 = ctx->d_AK[0]; a.AK1 = ctx->d_AK[1];
    a.Dsh = ctx->d_Dsh;
    a.tau = ctx->screening ? ctx->tau : 0.0;
    a.dmax = std::max((double)dmaxf, 1e-300);
    a.boys = ctx->d_boys;
    a.rank = ctx->rank; a.world = ctx->world;

    CK(cudaEventRecord(ctx->ev_fork, ms));
    for (int s = 0; s < ctx->nstreams; ++s) CK(cudaStreamWaitEvent(ctx->streams[s], ctx->ev_fork, 0));
    ctx->launches.clear();
    int nl = 0;
    const int ng = (int)ctx->groups.size();
    // Launch plan.  Kets per thread: long lists amortise the per-CTA prologue / J_ab reduction over many kets,
    // short lists are cut finer so that the grid still fills the 148 SMs (d-bra classes, per-rank share of a
    // multi-GPU run), and the serial work of one thread is capped so that highly contracted classes
    // (K = 36 x 36 primitive quartets per shell quartet) do not become a latency tail.  Order: the launches
    // whose threads run longest go first (they overlap with everything else), round-robin over the streams.
    struct Planned { int gi, gj, nbra, nket_max, kpt; double serial, cost; };
    std::vector<Planned> plan;
    for (int gi = ng - 1; gi >= 0; --gi)
        for (int gj = gi; gj >= 0; --gj) {
            const Group& bra = ctx->groups[gi];
            const Group& ket = ctx->groups[gj];
            if (a.tau > 0.0 && q_bucket_ceiling(bra.pairs[0].Q) * q_bucket_ceiling(ket.pairs[0].Q) * a.dmax < a.tau) continue;
            const int nbra = (bra.dev.npair - ctx->rank + ctx->world - 1) / ctx->world;
            if (nbra <= 0) continue;
            const int nket_max = gi == gj ? bra.dev.npair : ket.dev.npair;
            const double per_quartet = (double)bra.K * ket.K * model_flops_prim(bra.la, bra.lb, ket.la, ket.lb) + 400.0;
            const bool slab = bra.la == 2 && bra.lb >= 1;
            const int cta_threads = slab ? 128 : ctx->block;
            const long long want_chunks = (ctx->target_ctas + nbra - 1) / nbra;
            int kpt = (int)(nket_max / (want_chunks * cta_threads));
            kpt = std::min(kpt, (int)(ctx->serial_cap / per_quartet));
            kpt = std::max(1, std::min(kpt, ctx->kets_per_thread));
            const double nq = (double)bra.dev.npair * ket.dev.npair * (gi == gj ? 0.5 : 1.0);
            plan.push_back({gi, gj, nbra, nket_max, kpt, kpt * per_quartet, nq * per_quartet});
        }
    std::stable_sort(plan.begin(), plan.end(), [](const Planned& x, const Planned& y) {
        return x.serial != y.serial ? x.serial > y.serial : x.cost > y.cost;
    });
    for (const Planned& pl : plan) {
            const int gi = pl.gi, gj = pl.gj, nbra = pl.nbra, nket_max = pl.nket_max, kpt = pl.kpt;
            const Group& bra = ctx->groups[gi];
            const Group& ket = ctx->groups[gj];
            const ClassLaunch* cl = class_table(bra.cls, ket.cls);
            BuildArgs al = a;
            al.counter = ctx->d_counters + nl;
            if (ctx->profile) {
                while ((int)ctx->prof_ev.size() < 2 * (nl + 1)) { cudaEvent_t e; CK(cudaEventCreate(&e)); ctx->prof_ev.push_back(e); }
                CK(cudaEventRecord(ctx->prof_ev[2 * nl], ctx->streams[0]));
            }
            cl->jk(nk, nbra, nket_max, ctx->block, kpt, ctx->streams[ctx->profile ? 0 : (nl % ctx->nstreams)], bra.dev, ket.dev, al, gi == gj ? 1 : 0);
            if (ctx->profile) CK(cudaEventRecord(ctx->prof_ev[2 * nl + 1], ctx->streams[0]));
            ctx->launches.push_back({gi, gj});
            ++nl;
        }
    CK(cudaGetLastError());
    for (int s = 0; s < ctx->nstreams; ++s) {
        CK(cudaEventRecord(ctx->ev_join[s], ctx->streams[s]));
        CK(cudaStreamWaitEvent(ms, ctx->ev_join[s], 0));
    }
    finalize_kernel<<<nblk, tpb, 0, ms>>>(N, mode, ctx->d_fscale, ctx->d_AJ, ctx->d_AK[0], ctx->d_AK[1], dG0, dG1);
    CK(cudaEventRecord(ctx->ev_t1, ms));
    CK(cudaGetLastError());
    ctx->stats.launches = nl + 3;
    ctx->counters_pending = true;
    return QCF_OK;
}

int collect_stats(qcf_ctx* ctx) {
    if (!ctx->counters_pending) return QCF_OK;
    CK(cudaEventSynchronize(ctx->ev_t1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
    ctx->stats.kernel_ms = ms;
    const int nl = (int)ctx->launches.size();
    std::vector<unsigned long long> cnt(std::max(nl, 1));
    if (nl) CK(cudaMemcpy(cnt.data(), ctx->d_counters, sizeof(unsigned long long) * nl, cudaMemcpyDeviceToHost));
    long long q = 0;
    double flops = 0;
    const double digest = 12.0;  // per unique contracted integral (RHF model)
    ctx->launch_cnt.assign(cnt.begin(), cnt.begin() + nl);
    for (int i = 0; i < nl; ++i) {
        if (ctx->profile) CK(cudaEventElapsedTime(&ctx->launches[i].ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        const Group& b = ctx->groups[ctx->launches[i].bra];
        const Group& k = ctx->groups[ctx->launches[i].ket];
        q += (long long)cnt[i];
        const double nint = (double)ncart(b.la) * ncart(b.lb) * ncart(k.la) * ncart(k.lb);
        flops += (double)cnt[i] * ((double)b.K * k.K * model_flops_prim(b.la, b.lb, k.la, k.lb) + digest * nint);
    }
    ctx->stats.quartets = q;
    ctx->stats.model_flops = flops;
    ctx->counters_pending = false;
    return QCF_OK;
}

int host_build(qcf_ctx* ctx, int mode, const double* Pa, const double* Pb, double* G0, double* G1) {
    const size_t nn = (size_t)ctx->N * ctx->N;
    cudaStream_t ms = ctx->main_stream;
    cudaEvent_t e0 = ctx->ev_h0, e1 = ctx->ev_h1;
    CK(cudaEventRecord(e0, ms));
    std::memcpy(ctx->h_pin, Pa, nn * sizeof(double));
    CK(cudaMemcpyAsync(ctx->d_Pin[0], ctx->h_pin, nn * sizeof(double), cudaMemcpyHostToDevice, ms));
    if (Pb) {
        std::memcpy(ctx->h_pin + nn, Pb, nn * sizeof(double));
        CK(cudaMemcpyAsync(ctx->d_Pin[1], ctx->h_pin + nn, nn * sizeof(double), cudaMemcpyHostToDevice, ms));
    }
    int rc = run_build(ctx, mode, ctx->d_Pin[0], Pb ? ctx->d_Pin[1] : nullptr, ctx->d_G[0], ctx->d_G[1], ms);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_G[0], nn * sizeof(double), cudaMemcpyDeviceToHost, ms));
    if (G1) CK(cudaMemcpyAsync(ctx->h_pin + nn, ctx->d_G[1], nn * sizeof(double), cudaMemcpyDeviceToHost, ms));
    CK(cudaEventRecord(e1, ms));
    CK(cudaStreamSynchronize(ms));
    std::memcpy(G0, ctx->h_pin, nn * sizeof(double));
    if (G1) std::memcpy(G1, ctx->h_pin + nn, nn * sizeof(double));
    float t = 0;
    CK(cudaEventElapsedTime(&t, e0, e1));
    ctx->stats.total_ms = t;
    return collect_stats(ctx);
}

}  // namespace

// ======================================================================================================
extern "C" {

int qcf_create(const qcf_basis* b, const qcf_opts* o, qcf_ctx** out) {
    if (!b || !out) return QCF_ERR_ARG;
    *out = nullptr;
    qcf_ctx* ctx = new qcf_ctx();
    auto fail = [&](int code, const std::string& msg) {
        // keep the context alive so that qcf_last_error can report; caller destroys it
        ctx->err = msg;
        *out = ctx;
        return code;
    };
    if (b->cartesian != 1) return fail(QCF_ERR_ARG, "only Cartesian (6d) shells are supported");
    if (b->n_shells <= 0 || b->n_atoms <= 0) return fail(QCF_ERR_ARG, "empty basis");
    if (o) {
        ctx->device = o->device; ctx->rank = o->rank; ctx->world = o->world_size > 0 ? o->world_size : 1;
        if (o->block_threads > 0) ctx->block = o->block_threads;
        if (o->screen_tau < -0.5) ctx->screening = false;
        else if (o->screen_tau > 0) ctx->tau = o->screen_tau;
    }
    if (const char* e = getenv("QCF_PROFILE")) ctx->profile = (e[0] == '1');
    // measured on the N = 1007 build: 64 kets per thread is best when one GPU has the whole bra list, 32 for a
    // rank's share of it (finer chunks keep the smaller grids balanced)
    ctx->kets_per_thread = ctx->world > 1 ? 32 : 64;
    if (const char* e = getenv("QCF_KETS_PER_THREAD")) ctx->kets_per_thread = std::max(1, atoi(e));
    if (const char* e = getenv("QCF_STREAMS")) ctx->nstreams = std::min((int)qcf_ctx::MAXSTREAM, std::max(1, atoi(e)));
    if (const char* e = getenv("QCF_SERIAL_CAP")) ctx->serial_cap = std::max(1.0, atof(e));
    if (const char* e = getenv("QCF_TARGET_CTAS")) ctx->target_ctas = std::max(1, atoi(e));
    if (ctx->rank < 0 || ctx->rank >= ctx->world) return fail(QCF_ERR_ARG, "rank outside [0, world_size)");
    if (ctx->block < 32 || ctx->block > 128 || ctx->block % 32) return fail(QCF_ERR_ARG, "block_threads must be 32, 64, 96 or 128");
    ctx->natoms = b->n_atoms; ctx->nshell = b->n_shells;
    ctx->xyz.assign(b->xyz, b->xyz + 3 * b->n_atoms);
    ctx->charge.resize(b->n_atoms);
    for (int i = 0; i < b->n_atoms; ++i) ctx->charge[i] = b->Z ? (double)b->Z[i] : 0.0;
    ctx->sh_atom.assign(b->shell_atom, b->shell_atom + b->n_shells);
    ctx->sh_l.assign(b->shell_l, b->shell_l + b->n_shells);
    ctx->sh_np.assign(b->shell_nprim, b->shell_nprim + b->n_shells);
    ctx->sh_po.assign(b->shell_prim_off, b->shell_prim_off + b->n_shells);
    int nprim = 0;
    ctx->sh_off.assign(b->n_shells + 1, 0);
    for (int s = 0; s < b->n_shells; ++s) {
        if (ctx->sh_l[s] < 0 || ctx->sh_l[s] > LMAX) return fail(QCF_ERR_ARG, "angular momentum outside 0..2");
        if (ctx->sh_np[s] <= 0) return fail(QCF_ERR_ARG, "shell without primitives");
        if (ctx->sh_atom[s] < 0 || ctx->sh_atom[s] >= b->n_atoms) return fail(QCF_ERR_ARG, "shell_atom out of range");
        nprim = std::max(nprim, ctx->sh_po[s] + ctx->sh_np[s]);
        ctx->sh_off[s + 1] = ctx->sh_off[s] + ncart(ctx->sh_l[s]);
    }
    ctx->exps.assign(b->exps, b->exps + nprim);
    ctx->coefs.assign(b->coefs, b->coefs + nprim);
    ctx->N = ctx->sh_off.back();
    ctx->fscale.resize(ctx->N);
    for (int s = 0; s < b->n_shells; ++s) {
        const int l = ctx->sh_l[s];
        for (int i = 0; i < ncart(l); ++i) {
            // N(i,j,k)/N(l,0,0) = sqrt((2l-1)!! / ((2i-1)!!(2j-1)!!(2k-1)!!)); for l<=2 only xy,xz,yz differ: sqrt(3)
            const int x = cart_x(l, i), y = cart_y(l, i), z = cart_z(l, i);
            ctx->fscale[ctx->sh_off[s] + i] = (l == 2 && x < 2 && y < 2 && z < 2) ? std::sqrt(3.0) : 1.0;
        }
    }
    *out = ctx;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        ctx->err = "no CUDA device available (this engine has no CPU fallback)";
        return QCF_ERR_CUDA;
    }
    CK(cudaSetDevice(ctx->device));
    const size_t nn = (size_t)ctx->N * ctx->N;
    {
        std::vector<double> tab = make_boys_table();
        CK(upload(&ctx->d_boys, tab));
    }
    CK(upload(&ctx->d_fscale, ctx->fscale));
    CK(upload(&ctx->d_shoff, ctx->sh_off));
    CK(cudaStreamCreateWithFlags(&ctx->main_stream, cudaStreamNonBlocking));
    for (int s = 0; s < qcf_ctx::MAXSTREAM; ++s) {
        CK(cudaStreamCreateWithFlags(&ctx->streams[s], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_join[s], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreate(&ctx->ev_t0)); CK(cudaEventCreate(&ctx->ev_t1));
    CK(cudaEventCreate(&ctx->ev_h0)); CK(cudaEventCreate(&ctx->ev_h1));
    int rc = build_pairs(ctx);
    if (rc) return rc;
    for (int k = 0; k < 2; ++k) {
        CK(cudaMalloc(&ctx->d_Pin[k], nn * sizeof(double)));
        CK(cudaMalloc(&ctx->d_Pk[k], nn * sizeof(double)));
        CK(cudaMalloc(&ctx->d_AK[k], nn * sizeof(double)));
        CK(cudaMalloc(&ctx->d_G[k], nn * sizeof(double)));
    }
    CK(cudaMalloc(&ctx->d_Pj, nn * sizeof(double)));
    CK(cudaMalloc(&ctx->d_AJ, nn * sizeof(double)));
    CK(cudaMalloc(&ctx->d_Dsh, sizeof(float) * ctx->nshell * ctx->nshell));
    CK(cudaMalloc(&ctx->d_dmax, sizeof(unsigned int)));
    const int ng = (int)ctx->groups.size();
    ctx->max_launch = std::max(1, ng * (ng + 1) / 2);
    CK(cudaMalloc(&ctx->d_counters, sizeof(unsigned long long) * ctx->max_launch));
    CK(cudaMallocHost(&ctx->h_pin, 2 * nn * sizeof(double)));
    ctx->stats.n_basis = ctx->N; ctx->stats.n_shells = ctx->nshell;
    const long long nsp = (long long)ctx->nshell * (ctx->nshell + 1) / 2;
    ctx->stats.quartets_total = nsp * (nsp + 1) / 2;
    CK(cudaDeviceSynchronize());
    return QCF_OK;
}

int qcf_nbasis(const qcf_ctx* ctx) { return ctx ? ctx->N : QCF_ERR_ARG; }

int qcf_build_rhf(qcf_ctx* ctx, const double* P, double* G) {
    if (!ctx || !P || !G) return QCF_ERR_ARG;
    if (!ctx->d_AJ) { ctx->err = "context was not created successfully"; return QCF_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    return host_build(ctx, 0, P, nullptr, G, nullptr);
}

int qcf_build_uhf(qcf_ctx* ctx, const double* Pa, const double* Pb, double* Ga, double* Gb) {
    if (!ctx || !Pa || !Pb || !Ga || !Gb) return QCF_ERR_ARG;
    if (!ctx->d_AJ) { ctx->err = "context was not created successfully"; return QCF_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    return host_build(ctx, 1, Pa, Pb, Ga, Gb);
}

int qcf_build_jk(qcf_ctx* ctx, int nd, const double* const* P, double* const* J, double* const* K) {
    if (!ctx || nd <= 0 || !P || !J || !K) return QCF_ERR_ARG;
    if (!ctx->d_AJ) { ctx->err = "context was not created successfully"; return QCF_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    for (int d = 0; d < nd; ++d) {
        if (!P[d] || !J[d] || !K[d]) return QCF_ERR_ARG;
        int rc = host_build(ctx, 2, P[d], nullptr, J[d], K[d]);
        if (rc) return rc;
    }
    return QCF_OK;
}

int qcf_build_rhf_dev(qcf_ctx* ctx, const double* dP, double* dG, void* stream) {
    if (!ctx || !dP || !dG) return QCF_ERR_ARG;
    if (!ctx->d_AJ) { ctx->err = "context was not created successfully"; return QCF_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    return run_build(ctx, 0, dP, nullptr, dG, nullptr, (cudaStream_t)stream);
}

int qcf_build_uhf_dev(qcf_ctx* ctx, const double* dPa, const double* dPb, double* dGa, double* dGb, void* stream) {
    if (!ctx || !dPa || !dPb || !dGa || !dGb) return QCF_ERR_ARG;
    if (!ctx->d_AJ) { ctx->err = "context was not created successfully"; return QCF_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    return run_build(ctx, 1, dPa, dPb, dGa, dGb, (cudaStream_t)stream);
}

int qcf_eri_quartet(qcf_ctx* ctx, int s1, int s2, int s3, int s4, double* out) {
    if (!ctx || !out) return QCF_ERR_ARG;
    const int ns = ctx->nshell;
    if (s1 < 0 || s2 < 0 || s3 < 0 || s4 < 0 || s1 >= ns || s2 >= ns || s3 >= ns || s4 >= ns) return QCF_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    // canonical pairs: higher l first, ties by larger shell index first (as build_pairs does)
    auto canon = [&](int x, int y, bool& swapped) {
        int a = std::max(x, y), b = std::min(x, y);
        if (ctx->sh_l[b] > ctx->sh_l[a]) std::swap(a, b);
        swapped = (a != x);
        return std::make_pair(a, b);
    };
    bool sw12, sw34;
    auto p12 = canon(s1, s2, sw12), p34 = canon(s3, s4, sw34);
    const int n1 = ncart(ctx->sh_l[s1]), n2 = ncart(ctx->sh_l[s2]), n3 = ncart(ctx->sh_l[s3]), n4 = ncart(ctx->sh_l[s4]);
    const size_t ntot = (size_t)n1 * n2 * n3 * n4;
    auto i12 = ctx->pair_index.find(p12), i34 = ctx->pair_index.find(p34);
    if (i12 == ctx->pair_index.end() || i34 == ctx->pair_index.end()) {   // screened-out pair
        std::fill(out, out + ntot, 0.0);
        return QCF_OK;
    }
    const Group& g12 = ctx->groups[i12->second.first];
    const Group& g34 = ctx->groups[i34->second.first];
    const bool bra12 = g12.cls >= g34.cls;
    const Group& gb = bra12 ? g12 : g34;
    const Group& gk = bra12 ? g34 : g12;
    const int ib = bra12 ? i12->second.second : i34->second.second;
    const int ik = bra12 ? i34->second.second : i12->second.second;
    double* dout = nullptr;
    CK(cudaMalloc(&dout, ntot * sizeof(double)));
    class_table(gb.cls, gk.cls)->quartet(0, gb.dev, ib, gk.dev, ik, ctx->d_boys, dout);
    CK(cudaGetLastError());
    std::vector<double> tmp(ntot);
    CK(cudaMemcpy(tmp.data(), dout, ntot * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dout);
    // kernel layout: [bra a][bra b][ket c][ket d] with (a,b),(c,d) canonical; map back to (s1 s2|s3 s4)
    const int o1 = ctx->sh_off[s1], o2 = ctx->sh_off[s2], o3 = ctx->sh_off[s3], o4 = ctx->sh_off[s4];
    for (int i = 0; i < n1; ++i) for (int j = 0; j < n2; ++j) for (int k = 0; k < n3; ++k) for (int l = 0; l < n4; ++l) {
        // indices in canonical pair order
        const int a12 = sw12 ? j : i, b12 = sw12 ? i : j, na12 = sw12 ? n2 : n1, nb12 = sw12 ? n1 : n2;
        const int a34 = sw34 ? l : k, b34 = sw34 ? k : l, na34 = sw34 ? n4 : n3, nb34 = sw34 ? n3 : n4;
        (void)na12; (void)na34;
        size_t idx;
        if (bra12) idx = (((size_t)a12 * nb12 + b12) * (sw34 ? n4 : n3) + a34) * nb34 + b34;
        else idx = (((size_t)a34 * nb34 + b34) * (sw12 ? n2 : n1) + a12) * nb12 + b12;
        const double s = ctx->fscale[o1 + i] * ctx->fscale[o2 + j] * ctx->fscale[o3 + k] * ctx->fscale[o4 + l];
        out[(((size_t)i * n2 + j) * n3 + k) * n4 + l] = s * tmp[idx];
    }
    return QCF_OK;
}

int qcf_one_electron(qcf_ctx* ctx, double* S, double* T, double* V) {
    if (!ctx || !S || !T || !V) return QCF_ERR_ARG;
    if (!ctx->d_AJ) { ctx->err = "context was not created successfully"; return QCF_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    const size_t nn = (size_t)ctx->N * ctx->N;
    int *d_i = nullptr;
    double* d_d = nullptr;
    const int ns = ctx->nshell, na = ctx->natoms;
    std::vector<int> hi;
    hi.insert(hi.end(), ctx->sh_atom.begin(), ctx->sh_atom.end());
    hi.insert(hi.end(), ctx->sh_l.begin(), ctx->sh_l.end());
    hi.insert(hi.end(), ctx->sh_np.begin(), ctx->sh_np.end());
    hi.insert(hi.end(), ctx->sh_po.begin(), ctx->sh_po.end());
    std::vector<double> hd;
    hd.insert(hd.end(), ctx->exps.begin(), ctx->exps.end());
    hd.insert(hd.end(), ctx->coefs.begin(), ctx->coefs.end());
    hd.insert(hd.end(), ctx->xyz.begin(), ctx->xyz.end());
    hd.insert(hd.end(), ctx->charge.begin(), ctx->charge.end());
    CK(upload(&d_i, hi));
    CK(upload(&d_d, hd));
    double* d_out = nullptr;
    CK(cudaMalloc(&d_out, 3 * nn * sizeof(double)));
    CK(cudaMemset(d_out, 0, 3 * nn * sizeof(double)));
    ShellData sd{};
    sd.nshell = ns; sd.natoms = na; sd.N = ctx->N;
    sd.atom = d_i; sd.l = d_i + ns; sd.nprim = d_i + 2 * ns; sd.prim_off = d_i + 3 * ns; sd.off = ctx->d_shoff;
    const size_t npr = ctx->exps.size();
    sd.exps = d_d; sd.coefs = d_d + npr; sd.xyz = d_d + 2 * npr; sd.charge = d_d + 2 * npr + 3 * (size_t)na;
    sd.fscale = ctx->d_fscale;
    const long long npair = (long long)ns * (ns + 1) / 2;
    const long long nblk = (npair * 32 + 127) / 128;
    onee_kernel<<<(unsigned)nblk, 128>>>(sd, ctx->d_boys, d_out, d_out + nn, d_out + 2 * nn);
    CK(cudaGetLastError());
    CK(cudaMemcpy(S, d_out, nn * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(T, d_out + nn, nn * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(V, d_out + 2 * nn, nn * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_out); cudaFree(d_i); cudaFree(d_d);
    return QCF_OK;
}

int qcf_schwarz(qcf_ctx* ctx, double* Q) {
    if (!ctx || !Q) return QCF_ERR_ARG;
    const int ns = ctx->nshell;
    std::fill(Q, Q + (size_t)ns * ns, 0.0);
    for (const auto& g : ctx->groups)
        for (const auto& p : g.pairs) { Q[(size_t)p.sa * ns + p.sb] = p.Q; Q[(size_t)p.sb * ns + p.sa] = p.Q; }
    return QCF_OK;
}

int qcf_boys(qcf_ctx* ctx, int mmax, int n, const double* T, double* F) {
    if (!ctx || !T || !F || mmax < 0 || mmax > BOYS_LTOT || n <= 0) return QCF_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    double *dT = nullptr, *dF = nullptr;
    CK(cudaMalloc(&dT, sizeof(double) * n));
    CK(cudaMalloc(&dF, sizeof(double) * n * (mmax + 1)));
    CK(cudaMemcpy(dT, T, sizeof(double) * n, cudaMemcpyHostToDevice));
    const int g = (n + 127) / 128;
    switch (mmax) {
#define QCF_B(L) case L: boys_test_kernel<L><<<g, 128>>>(n, dT, ctx->d_boys, dF); break;
        QCF_B(0) QCF_B(1) QCF_B(2) QCF_B(3) QCF_B(4) QCF_B(5) QCF_B(6) QCF_B(7) QCF_B(8)
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(F, dF, sizeof(double) * n * (mmax + 1), cudaMemcpyDeviceToHost));
    cudaFree(dT); cudaFree(dF);
    return QCF_OK;
}

int qcf_fp64_peak(qcf_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return QCF_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    double* dout = nullptr;
    CK(cudaMalloc(&dout, sizeof(double)));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0, 0));
        fp64_peak_kernel<<<blocks, threads>>>(dout, iters, 1.0000001);
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dout);
    *tflops = best;
    return QCF_OK;
}

int qcf_stats(const qcf_ctx* cctx, qcf_stats_t* out) {
    if (!cctx || !out) return QCF_ERR_ARG;
    qcf_ctx* ctx = const_cast<qcf_ctx*>(cctx);
    int rc = collect_stats(ctx);
    if (rc) return rc;
    *out = ctx->stats;
    return QCF_OK;
}

int qcf_launch_profile(qcf_ctx* ctx, int max_rec, qcf_launch_rec* out) {
    if (!ctx || (max_rec > 0 && !out)) return QCF_ERR_ARG;
    int rc = collect_stats(ctx);
    if (rc) return rc;
    const int nl = (int)ctx->launches.size();
    for (int i = 0; i < nl && i < max_rec; ++i) {
        const Group& b = ctx->groups[ctx->launches[i].bra];
        const Group& k = ctx->groups[ctx->launches[i].ket];
        out[i] = {b.la, b.lb, b.K, k.la, k.lb, k.K, b.dev.npair, k.dev.npair,
                  i < (int)ctx->launch_cnt.size() ? (long long)ctx->launch_cnt[i] : 0,
                  model_flops_prim(b.la, b.lb, k.la, k.lb), ctx->launches[i].ms};
    }
    return nl;
}

const char* qcf_last_error(const qcf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void qcf_destroy(qcf_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& g : ctx->groups) free_group(g);
    cudaFree(ctx->d_boys); cudaFree(ctx->d_fscale); cudaFree(ctx->d_shoff);
    for (int k = 0; k < 2; ++k) { cudaFree(ctx->d_Pin[k]); cudaFree(ctx->d_Pk[k]); cudaFree(ctx->d_AK[k]); cudaFree(ctx->d_G[k]); }
    cudaFree(ctx->d_Pj); cudaFree(ctx->d_AJ); cudaFree(ctx->d_Dsh); cudaFree(ctx->d_dmax); cudaFree(ctx->d_counters);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    for (int s = 0; s < qcf_ctx::MAXSTREAM; ++s) { if (ctx->streams[s]) cudaStreamDestroy(ctx->streams[s]); if (ctx->ev_join[s]) cudaEventDestroy(ctx->ev_join[s]); }
    if (ctx->main_stream) cudaStreamDestroy(ctx->main_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->ev_h0) cudaEventDestroy(ctx->ev_h0);
    if (ctx->ev_h1) cudaEventDestroy(ctx->ev_h1);
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    delete ctx;
}

}  // extern "C"
