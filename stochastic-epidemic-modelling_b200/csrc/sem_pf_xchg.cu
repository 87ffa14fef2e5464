// sem_pf_xchg.cu -- one particle filter sharded over the GPUs of a node, exchange on the device.
//
// Every rank launches pf_persistent_x once per filter pass (cooperative launch: all of its CTAs are co-resident and may
// spin on each other and on the peers).  The body is the single-GPU whole-filter kernel (sem_pf_dev.cuh,
// pf_persistent_body<.., PUSH = true>): same SSA loops, same layouts, same weights; what changes is the resampling
// step, which globalises pmcmc.py:183-199 over peer memory -- see the "peer-memory exchange" block of sem_pf_dev.cuh.
// This file holds the kernel instantiations, the arena layout, CUDA IPC plumbing and the launch.
#include "sem_pf_host.h"

namespace sem {

template <class Model, int ARITH>
__global__ void __launch_bounds__(kMaxThreads) pf_persistent_x(const __grid_constant__ PfDev P, const __grid_constant__ XchgDev X) {
    pf_persistent_body<Model, ARITH, true>(P, &X);
}

template <class Model>
static const void *xchg_fn(int arith) {
    return arith == SEM_ARITH_UNIFORMIZED32 ? (const void *)pf_persistent_x<Model, SEM_ARITH_UNIFORMIZED32>
         : arith == SEM_ARITH_FAST32        ? (const void *)pf_persistent_x<Model, SEM_ARITH_FAST32>
                                            : nullptr;
}
static const void *xchg_kernel(const sem_pf_config *cfg) {
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1;
    switch (cfg->model) {
        case SEM_MODEL_SIR: return xchg_fn<SirModel>(cfg->arith);
#ifndef SEM_ONLY_SIR
        case SEM_MODEL_SEIR: return xchg_fn<SeirModel>(cfg->arith);
        default:
            switch (G) {
                case 1: return xchg_fn<SubModel<1>>(cfg->arith);
                case 2: return xchg_fn<SubModel<2>>(cfg->arith);
                case 3: return xchg_fn<SubModel<3>>(cfg->arith);
                default: return xchg_fn<SubModel<4>>(cfg->arith);
            }
#else
        default: return nullptr;
#endif
    }
}

struct XchgPlan { const void *fn; int threads, split_main, kper; size_t smem; };

static int xchg_plan(const sem_pf_config *cfg, int world, XchgPlan &pl) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (world < 1 || world > SEM_MAX_RANKS) { set_error("world must be 1..8"); return SEM_ERR_INVALID; }
    if (cfg->n_filters != 1 && world != 1) { set_error("the sharded filter runs one filter (n_filters = 1)"); return SEM_ERR_INVALID; }
    if (cfg->resampler != SEM_RESAMPLE_SYSTEMATIC) { set_error("the sharded filter resamples systematically"); return SEM_ERR_INVALID; }
    if (cfg->n_obs >= (1 << 20) - 1) { set_error("n_obs too large for the path-sampler token"); return SEM_ERR_INVALID; }
    if ((long long)cfg->n_particles * world > 0x7fffffffLL) { set_error("global particle count exceeds int32"); return SEM_ERR_INVALID; }
    pl.fn = xchg_kernel(cfg);
    if (!pl.fn) { set_error("the device-side exchange runs arith fast32 / uniformized32"); return SEM_ERR_INVALID; }
    const WsLayout w = ws_layout(cfg);
    pl.threads = persistent_threads(cfg, w, &pl.split_main);
    const int NB = w.nb * world;
    pl.kper = (NB + pl.threads - 1) / pl.threads;
    if (NB > 4096) { set_error("too many CTAs for the in-kernel combine (world * nb > 4096)"); return SEM_ERR_INVALID; }
    pl.smem = persistent_smem(cfg, NB, pl.threads, pl.split_main, true);
    int dev = 0, coop = 0, per_sm = 0;
    SEM_CUDA(cudaGetDevice(&dev));
    SEM_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) { set_error("device lacks cooperative launch"); return SEM_ERR_INVALID; }
    rc = persistent_prepare(pl.fn, pl.smem);
    if (rc) return rc;
    SEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pl.fn, pl.threads, pl.smem));
    if ((long long)per_sm * sm_count() < (long long)w.nb * cfg->n_filters) {
        set_error("this rank's particles are not co-resident on one GPU: use the host-driven exchange (sem_shard_*)");
        return SEM_ERR_INVALID;
    }
    return SEM_OK;
}

// Fill the exchange descriptor of one launch and launch the kernel.  arenas[r] = rank r's arena in this address space.
static int xchg_launch(const sem_pf_config *cfg, PfDev &P, const WsLayout &w, const XchgPlan &pl, const ArenaLayout &a, int world, int rank,
                       void *const *arenas, uint32_t generation, uint32_t tag, double timeout_s, bool want_iter, cudaStream_t s) {
    if (cfg->arith == SEM_ARITH_UNIFORMIZED32) {             // this translation unit's view of the candidate-count tables
        const int rcb = ktab_bind();
        if (rcb) return rcb;
    }
    P.j0 = rank * cfg->n_particles;
    P.split_main = pl.split_main;
    XchgDev X;
    memset(&X, 0, sizeof(X));
    X.W = world; X.rank = rank; X.NB = a.NB; X.kper = pl.kper;
    X.Ng = (long long)cfg->n_particles * world;
    X.gen0 = generation;
    X.tag = tag;
    int dev = 0;
    SEM_CUDA(cudaGetDevice(&dev));
    static int khz_of[64] = {0};                             // (cudaDevAttrClockRate is a slow query: once per device)
    if (dev >= 0 && dev < 64 && !khz_of[dev]) { int k = 0; SEM_CUDA(cudaDeviceGetAttribute(&k, cudaDevAttrClockRate, dev)); khz_of[dev] = k > 0 ? k : 1965000; }
    const int khz = (dev >= 0 && dev < 64) ? khz_of[dev] : 1965000;
    X.timeout = (long long)((timeout_s > 0 ? timeout_s : 20.0) * 1e3 * (khz > 0 ? khz : 1965000));
    for (int r = 0; r < world; r++) {
        char *base = (char *)arenas[r];
        X.part[r] = (double2 *)(base + a.part); X.rec[r] = (int32_t *)(base + a.rec);
        X.mail[r] = (unsigned long long *)(base + a.mail); X.iter[r] = (double *)(base + a.iter);
    }
    X.err = (int *)((char *)arenas[rank] + a.err);
    X.filter_stride = a.bytes;
    if (world == 1 && P.iter_out) X.iter[0] = P.iter_out;    // one rank: the caller's own result buffer
    else P.iter_out = want_iter ? X.iter[rank] : nullptr;
    void *args[] = {(void *)&P, (void *)&X};
    SEM_CUDA(cudaLaunchCooperativeKernel(pl.fn, dim3(w.nb, world == 1 ? cfg->n_filters : 1), dim3(pl.threads), args, pl.smem, s));
    return SEM_OK;
}

static int arena_mark_empty(const ArenaLayout &a, void *arena, cudaStream_t s) {
    char *base = (char *)arena;
    SEM_CUDA(cudaMemsetAsync(base, 0, a.part, s));                                  // error flag, mailbox
    SEM_CUDA(cudaMemsetAsync(base + a.part, 0xFF, a.iter - a.part, s));            // partial tables and records: all-ones = empty
    return SEM_OK;
}

bool xchg_single_available(const sem_pf_config *cfg) {
    XchgPlan pl;
    if (!push_eligible(cfg)) return false;
    const bool ok = xchg_plan(cfg, 1, pl) == SEM_OK;
    if (!ok) cudaGetLastError();
    return ok;
}

int xchg_run_single(const sem_pf_config *cfg, PfDev &P, const WsLayout &w, void *arena, cudaStream_t s, bool *launched) {
    *launched = false;
    XchgPlan pl;
    if (!push_eligible(cfg) || !w.xarena_bytes) return SEM_OK;
    if (xchg_plan(cfg, 1, pl) != SEM_OK) { cudaGetLastError(); return SEM_OK; }
    const ArenaLayout a = arena_layout(cfg, 1);
    // the workspace is scratch: empty marks per launch, one arena per filter (all-ones everywhere; the error flag counts as
    // set only when it is 1, and one rank never reads its mailbox)
    int rc = SEM_OK;
    SEM_CUDA(cudaMemsetAsync(arena, 0xFF, a.bytes * (size_t)cfg->n_filters, s));
    if (rc) return rc;
    void *arenas[1] = {arena};
    rc = xchg_launch(cfg, P, w, pl, a, 1, 0, arenas, 0u, 1u, 0.0, P.iter_out != nullptr, s);
    if (rc) return rc;
    *launched = true;
    return SEM_OK;
}

}  // namespace sem

using namespace sem;

extern "C" {

size_t sem_xchg_bytes(const sem_pf_config *cfg, int32_t world) {
    if (validate(cfg) || world < 1 || world > SEM_MAX_RANKS) return 0;
    return arena_layout(cfg, world).bytes;
}

int sem_xchg_alloc(size_t bytes, void **arena, unsigned char *ipc_handle) {
    if (!arena || !bytes) { set_error("bad arena request"); return SEM_ERR_INVALID; }
    SEM_CUDA(cudaMalloc(arena, bytes));
    if (ipc_handle) {
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, *arena);
        if (e != cudaSuccess) { cudaFree(*arena); *arena = nullptr; set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return SEM_ERR_CUDA; }
        static_assert(sizeof(h) == 64, "CUDA IPC handle size");
        memcpy(ipc_handle, &h, sizeof(h));
    }
    return SEM_OK;
}

int sem_xchg_open(const unsigned char *ipc_handle, void **peer_arena) {
    if (!ipc_handle || !peer_arena) { set_error("null"); return SEM_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    SEM_CUDA(cudaIpcOpenMemHandle(peer_arena, h, cudaIpcMemLazyEnablePeerAccess));
    return SEM_OK;
}

int sem_xchg_close(void *peer_arena) {
    if (peer_arena) SEM_CUDA(cudaIpcCloseMemHandle(peer_arena));
    return SEM_OK;
}

int sem_xchg_free(void *arena) {
    if (arena) SEM_CUDA(cudaFree(arena));
    return SEM_OK;
}

int sem_peer_enable(int32_t device, int32_t peer_device) {
    if (device == peer_device) return SEM_OK;
    int prev = 0, can = 0;
    SEM_CUDA(cudaGetDevice(&prev));
    SEM_CUDA(cudaDeviceCanAccessPeer(&can, device, peer_device));
    if (!can) { set_error("no peer access between the two devices"); return SEM_ERR_INVALID; }
    SEM_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    cudaSetDevice(prev);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return SEM_ERR_CUDA; }
    cudaGetLastError();
    return SEM_OK;
}

int sem_xchg_reset(const sem_pf_config *cfg, int32_t world, void *arena, void *stream) {
    if (validate(cfg) || !arena || world < 1 || world > SEM_MAX_RANKS) { set_error("bad arena reset"); return SEM_ERR_INVALID; }
    const ArenaLayout a = arena_layout(cfg, world);
    int rc = arena_mark_empty(a, arena, (cudaStream_t)stream);
    if (rc) return rc;
    SEM_CUDA(cudaMemsetAsync((char *)arena + a.iter, 0, a.bytes - a.iter, (cudaStream_t)stream));
    return SEM_OK;
}

double *sem_xchg_iteration_result(const sem_pf_config *cfg, int32_t world, void *arena) {
    if (validate(cfg) || !arena || world < 1 || world > SEM_MAX_RANKS) return nullptr;
    return (double *)((char *)arena + arena_layout(cfg, world).iter);
}

int sem_pf_sharded_supported(const sem_pf_config *cfg, int32_t world) {
    XchgPlan pl;
    return xchg_plan(cfg, world, pl) == SEM_OK ? 1 : 0;
}

int sem_pf_run_sharded(const sem_pf_config *cfg, const sem_pf_buffers *buf, sem_xchg_desc *x, void *stream) {
    if (!x) { set_error("null exchange descriptor"); return SEM_ERR_INVALID; }
    XchgPlan pl;
    int rc = xchg_plan(cfg, x->world, pl);
    if (rc) return rc;
    if (x->rank < 0 || x->rank >= x->world) { set_error("bad rank"); return SEM_ERR_INVALID; }
    for (int r = 0; r < x->world; r++) if (!x->arena[r]) { set_error("arena of a rank is not mapped"); return SEM_ERR_INVALID; }
    PfDev P; WsLayout w; bool replay;
    rc = fill_dev(cfg, buf, P, w, replay);
    if (rc) return rc;
    if (replay) { set_error("the sharded filter runs in Philox mode"); return SEM_ERR_INVALID; }
    const ArenaLayout a = arena_layout(cfg, x->world);
    if (buf->iteration_result && buf->iteration_result != (double *)((char *)x->arena[x->rank] + a.iter)) {
        set_error("iteration_result must be sem_xchg_iteration_result(arena)"); return SEM_ERR_INVALID;
    }
    const uint32_t tag = (x->launch_tag % 4095u) + 1u;
    rc = xchg_launch(cfg, P, w, pl, a, x->world, x->rank, x->arena, x->generation, tag, x->timeout_s, buf->iteration_result != nullptr,
                     (cudaStream_t)stream);
    if (rc) return rc;
    x->generation = (x->generation + (uint32_t)(cfg->n_obs > 1 ? cfg->n_obs - 1 : 0)) % 6u;   // only its value mod 2 and mod 3 matters
    x->launch_tag = tag;
    return SEM_OK;
}

// One MH iteration of the sharded filter enqueued by one call: H2D of theta from pinned host memory, this rank's launch
// with the path sample over the shards, D2H of the packed result (every rank receives the same one).  No synchronisation.
int sem_pf_iteration_sharded(const sem_pf_config *cfg, const sem_pf_buffers *buf, sem_xchg_desc *x, const double *theta_host,
                             double *result_host, void *stream) {
    if (!cfg || !buf || !buf->theta || !buf->iteration_result || !theta_host || !result_host) { set_error("sem_pf_iteration_sharded: null buffer"); return SEM_ERR_INVALID; }
    int rc = validate(cfg);
    if (rc) return rc;
    const int G = cfg->model >= SEM_MODEL_SIR_SUBGROUPS ? cfg->n_groups : 1, C = model_cols(cfg->model, G);
    cudaStream_t s = (cudaStream_t)stream;
    SEM_CUDA(cudaMemcpyAsync((void *)buf->theta, theta_host, model_ntheta(cfg->model, G) * sizeof(double), cudaMemcpyHostToDevice, s));
    rc = sem_pf_run_sharded(cfg, buf, x, stream);
    if (rc) return rc;
    SEM_CUDA(cudaMemcpyAsync(result_host, buf->iteration_result, (SEM_ITER_HEADER + (size_t)cfg->n_obs * C) * sizeof(double), cudaMemcpyDeviceToHost, s));
    return SEM_OK;
}

#ifdef SEM_PHASES
int sem_debug_phases_x(unsigned long long *host_out) {
    SEM_CUDA(cudaMemcpyFromSymbol(host_out, g_phase, sizeof(unsigned long long) * 24 * 256));
    return SEM_OK;
}
int sem_debug_cta_times_x(unsigned long long *host_out) {
    SEM_CUDA(cudaMemcpyFromSymbol(host_out, g_cta_t, sizeof(unsigned long long) * 2 * 128 * 160));
    return SEM_OK;
}
#endif

}  // extern "C"
