/* qcfock.h -- C ABI of the B200-native direct-SCF Fock-build engine for qchem-rs.
 *
 * This is the drop-in boundary.  The reference has no FFI today: the path is a plain
 * Rust call `molint::eri(&MolecularSystem) -> EriTensor` (core/src/hf/rhf.rs:45, uhf.rs:55) whose
 * result is repacked (rhf.rs:58-62) and contracted with the density once per SCF iteration
 * (rhf.rs:67-68 -> :152-167, uhf.rs:90-91 -> :210-227).  The engine fuses those pieces into one call
 * per iteration, G = fock_build(P), and never stores the N^4 tensor.
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on failure, never throws / unwinds
 *     across the boundary; qcf_last_error() gives the text of the last failure on that context;
 *   - matrices are dense column-major N x N float64 (nalgebra `DMatrix<f64>`); they are symmetric,
 *     so row- and column-major views coincide;
 *   - the caller owns all host buffers, the library copies on entry; the library owns device memory
 *     and streams; calls on one context are not re-entrant;
 *   - there is NO CPU fallback: without a usable CUDA device qcf_create fails with QCF_ERR_CUDA.
 */
#ifndef QCFOCK_H
#define QCFOCK_H

#ifdef __cplusplus
extern "C" {
#endif

#define QCF_OK 0
#define QCF_ERR_ARG (-1)      /* bad argument (null pointer, l > 2, spherical shells, ...)   */
#define QCF_ERR_CUDA (-2)     /* CUDA runtime failure; text in qcf_last_error              */
#define QCF_ERR_STATE (-3)    /* call not valid for this context                            */

typedef struct qcf_ctx qcf_ctx; /* opaque */

/* Molecule + basis in flat form.  Replaces what the reference reaches through
 * `&MolecularSystem` (atoms: rhf.rs:36, 116-117; shells: inside molint).  Fused SP shells are already
 * split; `coefs[k]` = contraction coefficient x normalisation of x^l exp(-a r^2); the engine applies
 * the remaining per-component factor (xy vs xx).  Basis functions are shell-major in the order given,
 * Cartesian components x,y,z / xx,xy,xz,yy,yz,zz.  Positions in bohr. */
typedef struct {
    int n_atoms;
    const int* Z;              /* [n_atoms]                                    */
    const double* xyz;         /* [3*n_atoms] bohr                             */
    int n_shells;
    const int* shell_atom;     /* [n_shells] index into the atom arrays        */
    const int* shell_l;        /* [n_shells] 0, 1 or 2                         */
    const int* shell_nprim;    /* [n_shells]                                   */
    const int* shell_prim_off; /* [n_shells] first primitive in exps/coefs     */
    const double* exps;
    const double* coefs;
    int cartesian;             /* must be 1 (6d)                               */
} qcf_basis;

/* ---- Loaders: native stand-ins for `BasisSet::load(path)` and
 * `MolecularSystem::load(path, &basis)` (qchem-cli/src/main.rs:76-77, 120-121).  Reads the reference's own
 * data/basis/*.json (MolSSI BSE schema) and data/mol/*.json and produces the flat qcf_basis above (fused SP
 * shells split, zero coefficients dropped, primitive normalisation folded in).  Pure host code.
 * On failure the returned object (if any) carries the message: qcf_system_error(). */
typedef struct qcf_system qcf_system;
int qcf_system_load(const char* basis_json_path, const char* molecule_json_path, qcf_system** out);
const qcf_basis* qcf_system_basis(const qcf_system* sys);       /* valid until qcf_system_free */
int qcf_system_n_electrons(const qcf_system* sys);              /* sum of Z (rhf.rs:36)        */
int qcf_system_n_basis(const qcf_system* sys);                  /* system.n_basis() (rhf.rs:37) */
double qcf_system_nuclear_repulsion(const qcf_system* sys);     /* rhf.rs:110-122              */
const char* qcf_system_error(const qcf_system* sys);
void qcf_system_free(qcf_system* sys);

typedef struct {
    double screen_tau; /* skip quartets with Q_ab Q_cd D_max < tau; <= 0 selects the default 1e-12;
                          QCF_TAU_NONE (any value < -0.5) disables screening entirely            */
    int device;        /* first CUDA device ordinal of this context                               */
    int rank;          /* multi-PROCESS use: this context evaluates share `rank` of `world_size`   */
    int world_size;    /*   of the cost-balanced bra split; the caller sums the partial matrices   */
    int block_threads; /* 0 = default                                                             */
    int n_gpus;        /* single-process multi-GPU: the context drives the devices  */
                       /*   device .. device+n_gpus-1 from the calling host thread, replicates P, */
                       /*   splits the bra list by modelled cost and sums the partial matrices     */
                       /*   over NVLink peer memory inside the finalize kernel; 0 or 1 = one GPU   */
    int deterministic; /* 1: fixed-point accumulation (64-bit integer atomics) and fixed-order     */
                       /*   reductions -- results are bitwise reproducible from run to run         */
} qcf_opts;
#define QCF_TAU_NONE (-1.0)

typedef struct {
    int n_basis, n_shells, n_pairs, n_groups;
    long long quartets;        /* unique shell quartets evaluated by the last build (this rank)   */
    long long quartets_total;  /* unique shell quartets before any screening (whole problem)      */
    double model_flops;        /* SURVEY.md 8d op-count model over the evaluated quartets          */
    double kernel_ms;          /* device time of the last build's kernels (CUDA events)           */
    double total_ms;           /* last build, including H2D / D2H copies                          */
    int launches;              /* kernels launched by the last build                              */
    long long prim_pairs;      /* primitive pairs of all shell pairs before primitive screening    */
    long long prim_pairs_kept; /* ... and those the kernels loop over                               */
    double create_ms;          /* wall time of qcf_create (pairs, Schwarz, upload, launch plan)     */
    double host_ms;            /* host time the last build spent enqueueing work (no device waits) */
    int n_devices;             /* GPUs driven by this context                                      */
    int graph_launches;        /* CUDA-graph launches of the last build (0: stream launches)       */
    double rank_imbalance;     /* max / mean modelled cost over the ranks of the bra split          */
} qcf_stats_t;

/* Build shell pairs, Schwarz bounds and the device-resident pair data.  (Replaces the one-off
 * molint::eri call, rhf.rs:45 / uhf.rs:55.) */
int qcf_create(const qcf_basis* basis, const qcf_opts* opts, qcf_ctx** out);
int qcf_nbasis(const qcf_ctx* ctx);

/* RHF: G = J[P] - K[P]/2, P carrying the factor 2 (rhf.rs:169-181).  Replaces rhf.rs:58-62 + :152-167. */
int qcf_build_rhf(qcf_ctx* ctx, const double* P, double* G);
/* UHF: Ga = J[Pa+Pb] - K[Pa], Gb = J[Pa+Pb] - K[Pb].  Replaces the two calls at uhf.rs:90-91 (:210-227). */
int qcf_build_uhf(qcf_ctx* ctx, const double* Pa, const double* Pb, double* Ga, double* Gb);
/* J_d = sum_kl P_d,kl (ij|kl),  K_d = sum_kl P_d,kl (ik|jl)  for nd densities. */
int qcf_build_jk(qcf_ctx* ctx, int nd, const double* const* P, double* const* J, double* const* K);

/* Incremental (difference-density) builds: G = G_prev + G(P - P_prev), screened on the block maxima of
 * P - P_prev, which shrink as the SCF converges.  The context keeps P_prev and G_prev on the device;
 * `reset` != 0 (or the first call) starts from P_prev = 0, i.e. a full build.  Same result as
 * qcf_build_rhf / qcf_build_uhf up to the screening threshold per build. */
int qcf_build_rhf_incremental(qcf_ctx* ctx, const double* P, double* G, int reset);
int qcf_build_uhf_incremental(qcf_ctx* ctx, const double* Pa, const double* Pb, double* Ga, double* Gb, int reset);

/* Device-resident variants: dP / dG are device pointers on the context's first device (N*N doubles
 * each); the call only ENQUEUES work on `stream` (a cudaStream_t of that device, may be 0) and returns
 * without waiting for the device.  With world_size > 1 the result is this rank's PARTIAL, already
 * symmetrised, matrix -- sum it over ranks with one allreduce (ncclAllReduce /
 * torch.distributed.all_reduce); with n_gpus > 1 it is the complete matrix. */
int qcf_build_rhf_dev(qcf_ctx* ctx, const double* dP, double* dG, void* stream);
int qcf_build_uhf_dev(qcf_ctx* ctx, const double* dPa, const double* dPb, double* dGa, double* dGb, void* stream);

/* One-electron matrices S (overlap), T (kinetic), V (nuclear attraction, charges Z), N x N each.
 * Replaces molint::overlap / kinetic / nuclear (rhf.rs:41-43, uhf.rs:52-54), so that a driver linked
 * against this library no longer needs the absent molint crate. */
int qcf_one_electron(qcf_ctx* ctx, double* S, double* T, double* V);

/* Contracted two-electron integrals of one shell quartet, out[na*nb*nc*nd] row-major (ab|cd),
 * fully normalised.  Computed by the same device code as the Fock build (parity tests). */
int qcf_eri_quartet(qcf_ctx* ctx, int sa, int sb, int sc, int sd, double* out);
/* Schwarz factors Q[sa*n_shells+sb] = sqrt(max|(ab|ab)|), 0 for pairs dropped at creation. */
int qcf_schwarz(qcf_ctx* ctx, double* Q);
/* Boys function F_0..F_mmax(T) as evaluated on the device (parity tests), n values of T. */
int qcf_boys(qcf_ctx* ctx, int mmax, int n, const double* T, double* F);

/* Measured FP64 FMA peak of the context's device in TFLOP/s (dependent-chain microbenchmark). */
int qcf_fp64_peak(qcf_ctx* ctx, double* tflops);

/* ---- Device-resident SCF iteration -------------------------------------------------
 * The per-iteration linear algebra of the reference's loops -- F = H + G, FPS - SPF, DIIS, X^T F X, the
 * symmetric eigensolve, the density update, energy and convergence test (rhf.rs:66-104, uhf.rs:79-189,
 * diis.rs:28-59, utils.rs:20-36) -- on the GPU (cuBLAS / cuSOLVER FP64), so that P and G never leave HBM
 * between iterations.  Same quirks as the reference loop; see qchem-rs_b200/csrc/scf_device.cu.
 *   qcf_scf_init : S, H = T + V (N x N, host); builds X = S^-1/2 and the Hueckel guess density on the device.
 *                  unrestricted = 0: RHF with n_alpha doubly occupied orbitals (n_beta ignored);
 *                  unrestricted = 1: UHF.  full_rebuild_every = 0: every build is a full build; k > 0:
 *                  difference-density (incremental) builds with a full rebuild every k-th iteration.
 *   qcf_scf_step : one iteration; `converged` is the reference's test (density_rms < epsilon).
 *   qcf_scf_get  : what = 0 density, 1 Fock matrix, 2 G (N*N doubles), 3 orbital energies (N doubles). */
typedef struct {
    int iteration;             /* loop index of this step (rhf.rs:66)                                */
    int converged;
    double electronic_energy;  /* 1/2 tr(P_new (2H + G(P_old)))  (rhf.rs:84-85)                      */
    double density_rms;        /* diagonal-only rms of the density change (rhf.rs:87-88; uhf.rs:137) */
    double build_ms;           /* device time of the Fock build                                      */
    double linalg_ms;          /* device time of everything else in the step                         */
    double wall_ms;            /* host wall time of the step                                         */
} qcf_scf_info;
int qcf_scf_init(qcf_ctx* ctx, const double* S, const double* H, int unrestricted, int n_alpha, int n_beta,
                 int full_rebuild_every);
int qcf_scf_step(qcf_ctx* ctx, double epsilon, qcf_scf_info* out);
int qcf_scf_get(qcf_ctx* ctx, int what, int spin, double* out);

int qcf_stats(const qcf_ctx* ctx, qcf_stats_t* out);
/* Device time (ms, CUDA events) each GPU of the context spent on its share of the last build; returns the
 * number of devices (writes at most max_dev values). */
int qcf_device_times(qcf_ctx* ctx, int max_dev, double* ms);

/* Per-launch record of the last build: class (la lb|lc ld), primitive-pair counts, list lengths,
 * unique shell quartets evaluated, model flops per primitive quartet, and -- only when the context was
 * created with the environment variable QCF_PROFILE=1, which serialises the launches -- its device time.
 * Returns the number of launches of the last build (writes at most max_rec records). */
typedef struct {
    int la, lb, kab, lc, ld, kcd, nbra, nket;
    long long quartets;
    double flops_per_prim_quartet;
    float ms;
} qcf_launch_rec;
int qcf_launch_profile(qcf_ctx* ctx, int max_rec, qcf_launch_rec* out);
const char* qcf_last_error(const qcf_ctx* ctx);
void qcf_destroy(qcf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* QCFOCK_H */
