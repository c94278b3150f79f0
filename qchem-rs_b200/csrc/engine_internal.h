// engine_internal.h -- host-side state of the Fock-build engine, shared by engine.cu (pair data, builds, C ABI)
// and scf_device.cu (the device-resident SCF iteration).  Not part of the public boundary (include/qcfock.h).
#pragma once
#include "../../include/qcfock.h"
#include "eri_device.cuh"

#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <vector>

struct HostPair {
    int sa, sb;       // shell ids, shell sa has l >= shell sb
    double Q;
    int keff = -1;                      // primitive pairs kept (-1: all K)
    std::vector<uint16_t> order;        // primitive-pair order (by primitive Schwarz factor); empty: natural
};

// host description of one (la, lb, K) pair group; the device copies live in qcf_device::groups
struct HostGroup {
    int la, lb, K, cls;
    std::vector<HostPair> pairs;
    size_t pair_off = 0;                // offset of this group's pairs in the concatenated per-pair arrays
};

struct GroupDev {
    int *fa = nullptr, *fb = nullptr, *sa = nullptr, *sb = nullptr, *np = nullptr;
    double *Q = nullptr, *Qb = nullptr, *prim = nullptr, *AB = nullptr;
    int* bra_list = nullptr;            // this device's share of the group's bra pairs (null: all of them)
    int nbra = 0;
    qcf::PairGroup pg{};
};

struct PlannedLaunch {
    int gi, gj, kpt, ps;
    double serial, cost;
};

constexpr int QCF_STAGE_CHUNKS = 4;       // host-buffer calls move P and G in this many pipelined pieces
constexpr int QCF_MAXSTREAM = 32;
constexpr int QCF_MAXDEV = 16;

// everything that lives on one GPU
struct qcf_device {
    int device = 0;                     // CUDA ordinal
    int rank = 0;                       // global rank of this device in the bra split
    std::vector<GroupDev> groups;
    double* boys = nullptr;
    double* fscale = nullptr;
    int* shoff = nullptr;
    int *all_sa = nullptr, *all_sb = nullptr;   // concatenated shell ids of all pairs (pair_dmax_kernel)
    float* all_Dp = nullptr;                    // concatenated per-pair density maxima
    double *Pin[2] = {nullptr, nullptr}, *Pj = nullptr, *Pk[2] = {nullptr, nullptr};
    double *AJ = nullptr, *AK[2] = {nullptr, nullptr}, *G[2] = {nullptr, nullptr};
    double *Pprev[2] = {nullptr, nullptr}, *Gprev[2] = {nullptr, nullptr};   // incremental builds (allocated on first use)
    float* Dsh = nullptr;
    qcf::BuildScalars* sc = nullptr;
    unsigned long long* counters = nullptr;
    cudaStream_t main = nullptr;
    cudaStream_t streams[QCF_MAXSTREAM] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[QCF_MAXSTREAM] = {};
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;   // around this device's share of a build (timing)
    cudaEvent_t ev_acc = nullptr, ev_red = nullptr; // accumulators complete / cross-device reduction done
    // captured build: scale -> maxima -> memsets -> every class launch, keyed on (mode, input pointers)
    cudaGraphExec_t graph = nullptr;
    int graph_mode = -1;
    const double *graph_pa = nullptr, *graph_pb = nullptr;
    std::vector<int*> launch_list;      // per planned launch: this device's bra indices (per-launch split only)
    std::vector<int> launch_nbra;
    std::vector<cudaEvent_t> prof_ev;
    int launches = 0;
    float last_ms = 0;
};

struct qcf_ctx {
    std::string err;
    int block = 64, kets_per_thread = 64, target_ctas = 296, nstreams = 8;
    double serial_cap = 4e6;          // model flops one thread may run serially in one launch
    double tau = 1e-12;
    double red_eps_factor = 0.1;      // scatter contributions below this fraction of tau are not sent to memory
    double pair_cut_factor = 1e-2;    // shell pairs with Q Q_max below this fraction of tau are dropped at creation
    double prim_cut_factor = 1.0;     // primitive pairs with Q_k Q_max below this multiple of tau are dropped at creation
    bool screening = true, deterministic = false, use_graph = true, profile = false;
    int stream_prio = 0;              // 0: one priority; 1: block kernels on high-priority streams, slab on low; 2: reverse (QCF_PRIO)
    int launch_order = 0;             // 0: longest-running threads first; 1: biggest launches first (QCF_ORDER)
    int ps_min_prim = 36;             // primitive quartets per shell quartet from which lanes share a quartet
    int world = 1;                    // total number of ranks in the bra split (processes x devices)
    // basis (host copies)
    int natoms = 0, nshell = 0, N = 0;
    std::vector<double> xyz, exps, coefs, charge;
    std::vector<int> sh_atom, sh_l, sh_np, sh_po, sh_off;
    std::vector<double> fscale;       // per basis function component scale
    std::vector<HostGroup> groups;
    std::map<std::pair<int, int>, std::pair<int, int>> pair_index;  // (sa,sb) canonical -> (group, index)
    std::vector<PlannedLaunch> plan;
    std::vector<std::vector<std::vector<int>>> bra_split;   // [group][rank] -> bra indices (cost-balanced)
    std::vector<double> rank_cost;                          // modelled cost per rank
    // Per-LAUNCH split (split_min_bras > 0, world > 1): a launch whose bra list is too short to fill `world` GPUs is
    // shared by fewer ranks (at least split_min_bras bra pairs each), so that the ranks run fewer, larger grids; the
    // long launches are split over all ranks and even out the load.  [plan index][rank] -> bra indices.
    std::vector<std::vector<std::vector<int>>> launch_split;
    int split_min_bras = 0;
    double qmax = 0;
    long long prim_total = 0, prim_kept = 0;
    size_t npairs = 0;
    std::vector<qcf_device> dev;      // dev[0] is the context's primary device
    double* h_pin = nullptr;          // pinned staging, 4*N*N
    struct qcf_copy_pool* pool = nullptr;   // helper threads of the host staging copies
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_h0 = nullptr, ev_h1 = nullptr, ev_in = nullptr;
    cudaEvent_t ev_out[2 * QCF_STAGE_CHUNKS] = {};   // one per D2H chunk of the host-buffer calls (pipelined staging)
    // stats of the last build
    struct LaunchRec { int bra, ket; float ms = 0; };
    std::vector<LaunchRec> launches;
    std::vector<unsigned long long> launch_cnt;
    qcf_stats_t stats{};
    bool counters_pending = false;
    int incremental_builds = 0;       // builds since the last full rebuild (incremental mode)
    struct qcf_scf* scf = nullptr;    // device-resident SCF state (scf_device.cu)
};

#define QCF_CK(call)                                                                                \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                         \
            return QCF_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

namespace qcf_internal {
// Enqueue one complete build on the context's devices.  mode 0 RHF, 1 UHF, 2 J/K.  dPa/dPb and dG0/dG1 are device
// pointers on dev[0]; the work is asynchronous on `user` (a stream of dev[0]) -- no host synchronisation.
int run_build(qcf_ctx* ctx, int mode, const double* dPa, const double* dPb, double* dG0, double* dG1, cudaStream_t user);
// the SCF step's build: `incremental` adds G(P - P_prev) to dG0/dG1 (which then must hold G(P_prev)); `reset` zeroes
// P_prev and dG first (full build through the same path)
int run_build_scf(qcf_ctx* ctx, int mode, const double* dPa, const double* dPb, double* dG0, double* dG1, cudaStream_t user,
                  bool incremental, bool reset);
int collect_stats(qcf_ctx* ctx);
void scf_destroy(qcf_ctx* ctx);
}  // namespace qcf_internal
