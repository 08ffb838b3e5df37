// Device-resident affine-invariant ensemble sampler: emcee's default red/blue StretchMove(a = 2)
// as the reference drives it (analysis/runner.py:403,416-419), with positions, log-probabilities,
// random numbers, proposals and accept/reject all on the GPU.  One emcee iteration is
//     split (random red/blue partition) -> lnprob kernel with fused propose/accept x 2 -> store
// (the likelihood kernel draws the proposals of the active half in its prologue and its finishing
// CTAs accept or reject in place: csrc/mcd_kernels.cu, FUSE) captured once as a CUDA graph and
// replayed per step: no host round trip inside a chain.
//
// emcee is a third-party dependency of the reference (unpinned, not vendored); the algorithm
// restated here is its published one (Goodman & Weare 2010; emcee 3 `RedBlueMove.propose`):
//   z = ((a-1) u + 1)^2 / a,  q = c_j - (c_j - s) z,  accept if (P-1) ln z + lnp(q) - lnp(s) > ln u'
// with the complementary walker j drawn uniformly and the second half seeing the updated first.
// Random numbers: Philox4x32-10, counter = (step, half, walker, stream id), key = seed.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <new>

#include "mcd_internal.h"
#include "mcd_rng.cuh"

using namespace mcd;

namespace {

constexpr int kMaxWalkers = 4096;   // split_kernel keeps one 64-bit key per walker in 32 KiB of shared memory

struct Ensemble {
    int n_segments;                     // independent ensembles advanced together (radial bins)
    int n_walkers, n_theta, n0, n1;     // per segment; n0 = first ("red") half, n1 = second
    uint64_t seed;
    double a;
    double *pos;          // [S][W][P]
    double *lnp;          // [S][W]
    double *lnp_q;        // [S][n0] scratch of the geometry dry run
    int *perm;            // [S][W]: perm[s][0:n0] = red walkers of segment s, perm[s][n0:W] = blue
    long long *n_accepted;   // [S][W]
    unsigned int *step;      // [2]: global step counter, step inside the current run() chunk
    double *chain;        // [chunk][S][W][P] or nullptr
    double *chain_lnp;    // [chunk][S][W]
};

// random red/blue partition: rank of a random key (emcee: inds = arange(W) % 2; shuffle(inds));
// one CTA per segment
__global__ void split_kernel(Ensemble E) {
    extern __shared__ unsigned long long keys[];
    const uint32_t step = E.step[0];
    const int seg = blockIdx.x;
    for (int w = threadIdx.x; w < E.n_walkers; w += blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4(step, 2u, (uint32_t)(seg * E.n_walkers + w), 7u),
                                      make_uint2((uint32_t)E.seed, (uint32_t)(E.seed >> 32)));
        keys[w] = (((unsigned long long)r.x << 32) | r.y);
    }
    __syncthreads();
    for (int w = threadIdx.x; w < E.n_walkers; w += blockDim.x) {
        const unsigned long long mine = keys[w];
        int rank = 0;
        for (int o = 0; o < E.n_walkers; ++o) {
            const unsigned long long other = keys[o];
            rank += (other < mine) || (other == mine && o < w);
        }
        E.perm[(size_t)seg * E.n_walkers + rank] = w;
    }
}

// end of a step: append the state to the chain (if one is kept) and advance the step counters.
// One CTA, so that the counters can be advanced after every thread has read them.
__global__ void __launch_bounds__(1024) store_advance_kernel(Ensemble E) {
    const unsigned int local = E.step[1];
    const int rows = E.n_walkers * E.n_segments;
    const int total = rows * E.n_theta;
    if (E.chain) {
        for (int i = threadIdx.x; i < total; i += blockDim.x) E.chain[(size_t)local * total + i] = E.pos[i];
        for (int i = threadIdx.x; i < rows; i += blockDim.x) E.chain_lnp[(size_t)local * rows + i] = E.lnp[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        E.step[0] += 1u;
        E.step[1] = local + 1u;
    }
}

}  // namespace

struct mcd_ensemble {
    mcd_handle *h = nullptr;
    int device = 0;
    Ensemble E{};
    cudaStream_t stream = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool graph_stores = false;
    size_t chain_cap_steps = 0;
    bool have_state = false;
    void *state = nullptr;                     // one stream-ordered allocation holding every array of E but the chain
    unsigned long long graph_generation = 0;   // handle_generation() the graph was captured at
    unsigned long long tag_base = 0;           // exchange tag of this ensemble's fused half-steps
    unsigned int steps_done = 0;     // host mirror of E.step[0]
    int last_path = 0;               // 1: resident-chain kernel, 2: CUDA graph of launches (diagnostics)
};

static void free_ensemble(mcd_ensemble *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->exec) cudaGraphExecDestroy(e->exec);
    if (e->graph) cudaGraphDestroy(e->graph);
    if (e->stream) {
        // stream-ordered frees into the device's memory pool: cudaFree synchronises the whole device and was
        // measured at 30-130 ms per call in a process that also holds a large catalogue (profiles/r02_summary.md)
        if (e->state) cudaFreeAsync(e->state, e->stream);
        if (e->E.chain) cudaFreeAsync(e->E.chain, e->stream);
        cudaStreamSynchronize(e->stream);
        forget_stream(e->h, e->stream);
        cudaStreamDestroy(e->stream);
    }
    delete e;
}

extern "C" void mcd_ensemble_destroy(mcd_ensemble *e) { free_ensemble(e); }

#define ENS_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t err__ = (call);                                                                      \
        if (err__ != cudaSuccess) return set_error(-2, "%s failed: %s", #call, cudaGetErrorString(err__)); \
    } while (0)

extern "C" int mcd_ensemble_create(mcd_handle *h, int32_t n_walkers, uint64_t seed, double stretch_a, mcd_ensemble **out) {
    if (!h || !out) return set_error(-1, "null argument");
    *out = nullptr;
    mcd_info info;
    if (mcd_get_info(h, &info) != 0) return -1;
    // emcee: "nwalkers >= 2 * ndim" (RuntimeError otherwise)
    if (n_walkers < 2 || n_walkers > kMaxWalkers || n_walkers < 2 * info.n_theta)
        return set_error(-1, "n_walkers = %d must be in [max(2, 2 * n_theta = %d), %d]", n_walkers, 2 * info.n_theta, kMaxWalkers);
    if (!(stretch_a > 1.0)) return set_error(-1, "the stretch scale a must exceed 1");
    mcd_ensemble *e = new (std::nothrow) mcd_ensemble();
    if (!e) return set_error(-4, "out of host memory");
    e->h = h;
    e->device = handle_device(h);
    // bits 36..61: ensembles created on this handle so far (every rank of a sharded run creates them in the
    // same order); bits 0..35: 2 * step + half
    e->tag_base = (1ull << 62) | ((next_fuse_nonce(h) & 0x3ffffffull) << 36);
    Ensemble &E = e->E;
    E.n_segments = std::max(1, (int)info.n_segments);
    E.n_walkers = n_walkers;
    E.n_theta = info.n_theta;
    E.n0 = (n_walkers + 1) / 2;
    E.n1 = n_walkers - E.n0;
    E.seed = seed;
    E.a = stretch_a;
    const size_t P = (size_t)std::max(1, E.n_theta);
    const size_t S = (size_t)E.n_segments;
    bool ok = cudaSetDevice(e->device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess;
    // one allocation from the device's memory pool (cudaMallocAsync: no device-wide synchronisation), carved up
    // at 256-byte boundaries
    auto round = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_pos = round(sizeof(double) * S * n_walkers * P), b_lnp = round(sizeof(double) * S * n_walkers);
    const size_t b_q = round(sizeof(double) * S * E.n0), b_perm = round(sizeof(int) * S * n_walkers);
    const size_t b_acc = round(sizeof(long long) * S * n_walkers), b_step = round(sizeof(unsigned int) * 2);
    const size_t total = b_pos + b_lnp + b_q + b_perm + b_acc + b_step;
    ok = ok && cudaMallocAsync(&e->state, total, e->stream) == cudaSuccess;
    if (ok) {
        char *base = static_cast<char *>(e->state);
        E.pos = reinterpret_cast<double *>(base);
        E.lnp = reinterpret_cast<double *>(base + b_pos);
        E.lnp_q = reinterpret_cast<double *>(base + b_pos + b_lnp);
        E.perm = reinterpret_cast<int *>(base + b_pos + b_lnp + b_q);
        E.n_accepted = reinterpret_cast<long long *>(base + b_pos + b_lnp + b_q + b_perm);
        E.step = reinterpret_cast<unsigned int *>(base + b_pos + b_lnp + b_q + b_perm + b_acc);
        ok = cudaMemsetAsync(E.n_accepted, 0, b_acc + b_step, e->stream) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(e->stream) == cudaSuccess;
    }
    if (!ok) {
        const cudaError_t err = cudaGetLastError();
        free_ensemble(e);
        return set_error(-2, "allocating the ensemble state failed: %s", cudaGetErrorString(err));
    }
    *out = e;
    return 0;
}

extern "C" int mcd_ensemble_set_state(mcd_ensemble *e, const double *pos_host) {
    if (!e || !pos_host) return set_error(-1, "null argument");
    ENS_CUDA(cudaSetDevice(e->device));
    Ensemble &E = e->E;
    ENS_CUDA(cudaMemcpyAsync(E.pos, pos_host, sizeof(double) * E.n_segments * E.n_walkers * E.n_theta, cudaMemcpyHostToDevice,
                             e->stream));
    if (int rc = launch_ensemble(e->h, E.pos, E.n_walkers, E.lnp, 1, e->stream)) return rc;
    ENS_CUDA(cudaStreamSynchronize(e->stream));
    e->have_state = true;
    return 0;
}

static int enqueue_step(mcd_ensemble *e) {
    Ensemble &E = e->E;
    const int split_threads = std::min(1024, ((E.n_walkers + 31) / 32) * 32);
    split_kernel<<<E.n_segments, split_threads, sizeof(unsigned long long) * E.n_walkers, e->stream>>>(E);
    for (int half = 0; half < 2; ++half) {
        const int ns = half == 0 ? E.n0 : E.n1;
        if (ns == 0) continue;
        // one launch per half-step: the likelihood kernel draws the proposals of the active half and
        // its finishing CTAs accept or reject them in place
        FuseParams f{};
        f.half = half;
        f.n0 = E.n0;
        f.walkers_total = E.n_walkers;
        f.a = E.a;
        f.seed = E.seed;
        f.pos = E.pos;
        f.lnp = E.lnp;
        f.perm = E.perm;
        f.n_accepted = E.n_accepted;
        f.step = E.step;
        f.tag_base = e->tag_base;
        if (int rc = launch_ensemble_fused(e->h, ns, f, e->stream)) return rc;
    }
    const int total = E.n_segments * E.n_walkers * std::max(1, E.n_theta);
    store_advance_kernel<<<1, std::min(1024, ((total + 31) / 32) * 32), 0, e->stream>>>(E);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : set_error(-2, "launching the ensemble step failed: %s", cudaGetErrorString(err));
}

static int build_graph(mcd_ensemble *e) {
    if (e->exec) {
        cudaGraphExecDestroy(e->exec);
        e->exec = nullptr;
    }
    if (e->graph) {
        cudaGraphDestroy(e->graph);
        e->graph = nullptr;
    }
    // size the likelihood scratch for both half-ensemble shapes outside the capture
    Ensemble &E = e->E;
    if (int rc = launch_ensemble(e->h, E.pos, E.n0, E.lnp_q, 1, e->stream)) return rc;
    if (E.n1 > 0)
        if (int rc = launch_ensemble(e->h, E.pos, E.n1, E.lnp_q, 1, e->stream)) return rc;
    ENS_CUDA(cudaStreamSynchronize(e->stream));
    ENS_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_step(e);
    cudaGraph_t g = nullptr;
    const cudaError_t end = cudaStreamEndCapture(e->stream, &g);
    if (rc != 0 || end != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        return rc ? rc : -2;
    }
    e->graph = g;
    ENS_CUDA(cudaGraphInstantiate(&e->exec, e->graph, 0));
    e->graph_generation = handle_generation(e->h);
    return 0;
}

// MCD_TIMING=1: wall-clock of the phases of a run on stderr (diagnostics)
static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec;
}

extern "C" int mcd_ensemble_run(mcd_ensemble *e, int32_t n_steps, double *chain_host, double *lnprob_host,
                                int64_t *n_accepted_host) {
    if (!e || n_steps < 0) return set_error(-1, "bad argument");
    const char *timing_env = getenv("MCD_TIMING");
    const bool timing = timing_env && timing_env[0] == '1';
    const double t_enter = now_ms();
    if (!e->have_state) return set_error(-1, "mcd_ensemble_set_state has not been called");
    ENS_CUDA(cudaSetDevice(e->device));
    Ensemble &E = e->E;
    const size_t P = (size_t)std::max(1, E.n_theta);
    const size_t rows = (size_t)E.n_walkers * E.n_segments;
    const size_t per_step = rows * P;
    const bool store = chain_host != nullptr || lnprob_host != nullptr;
    // chain chunks of at most 256 MiB on the device
    size_t chunk = std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, n_steps), ((size_t)256 << 20) / (per_step * 8)));
    if (store && chunk > e->chain_cap_steps) {
        // chain [chunk][rows][P] followed by lnprob [chunk][rows], one stream-ordered allocation
        if (E.chain) ENS_CUDA(cudaFreeAsync(E.chain, e->stream));
        E.chain = E.chain_lnp = nullptr;
        e->chain_cap_steps = 0;
        ENS_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&E.chain), sizeof(double) * chunk * (per_step + rows), e->stream));
        E.chain_lnp = E.chain + chunk * per_step;
        e->chain_cap_steps = chunk;
        if (e->exec) {   // pointers baked into the graph changed
            cudaGraphExecDestroy(e->exec);
            e->exec = nullptr;
        }
    }
    if (store) chunk = std::min(chunk, e->chain_cap_steps);

    // ---- small catalogues: whole chains inside one CTA per segment (csrc/mcd_kernels.cu) -----------
    const char *no_resident = getenv("MCD_NO_RESIDENT_CHAIN");
    if (!(no_resident && no_resident[0] == '1')) {
        int rc = 0;
        int done = 0;
        while (done < n_steps) {
            const int todo = (int)std::min<size_t>(chunk, (size_t)(n_steps - done));
            ChainParams c{};
            c.n_steps = todo;
            c.n_walkers = E.n_walkers;
            c.n0 = E.n0;
            c.step0 = e->steps_done;
            c.a = E.a;
            c.seed = E.seed;
            c.pos = E.pos;
            c.lnp = E.lnp;
            c.n_accepted = E.n_accepted;
            c.chain = store ? E.chain : nullptr;
            c.chain_lnp = store ? E.chain_lnp : nullptr;
            rc = launch_resident_chain(e->h, c, e->stream);
            if (rc != 0) break;              // 1: not eligible (nothing was launched), < 0: error
            if (chain_host) ENS_CUDA(cudaMemcpyAsync(chain_host + (size_t)done * per_step, E.chain,
                                                     sizeof(double) * todo * per_step, cudaMemcpyDeviceToHost, e->stream));
            if (lnprob_host) ENS_CUDA(cudaMemcpyAsync(lnprob_host + (size_t)done * rows, E.chain_lnp,
                                                      sizeof(double) * todo * rows, cudaMemcpyDeviceToHost, e->stream));
            ENS_CUDA(cudaStreamSynchronize(e->stream));
            rc = resident_chain_status(e->h, e->stream);
            if (rc != 0) break;
            e->steps_done += (unsigned int)todo;
            done += todo;
        }
        if (rc < 0) return rc;
        if (rc == 0) {
            e->last_path = 1;
            ENS_CUDA(cudaMemcpy(E.step, &e->steps_done, sizeof(unsigned int), cudaMemcpyHostToDevice));
            if (n_accepted_host)
                ENS_CUDA(cudaMemcpy(n_accepted_host, E.n_accepted, sizeof(long long) * rows, cudaMemcpyDeviceToHost));
            return 0;
        }
        // rc == 1 on the first chunk: fall through to the graph of launches
    }
    e->last_path = 2;
    // the graph bakes in whether the chain is stored (E.chain pointer): rebuild when that changes
    double *saved_chain = E.chain, *saved_lnp = E.chain_lnp;
    if (!store) E.chain = E.chain_lnp = nullptr;
    // the graph bakes in the handle's scratch buffers, packed columns, routing and exchange pointers: a
    // larger lnprob call, a re-pack or an exchange attach since the capture makes it stale
    const double t_graph = now_ms();
    if (!e->exec || e->graph_stores != store || e->graph_generation != handle_generation(e->h)) {
        const int rc = build_graph(e);
        if (rc) {
            E.chain = saved_chain;
            E.chain_lnp = saved_lnp;
            return rc;
        }
        e->graph_stores = store;
    }
    const double t_loop = now_ms();
    double t_queued = t_loop;
    int rc = order_on_stream(e->h, e->stream);
    for (int done = 0; done < n_steps && rc == 0;) {
        const int todo = (int)std::min<size_t>(chunk, (size_t)(n_steps - done));
        if (cudaMemsetAsync(E.step + 1, 0, sizeof(unsigned int), e->stream) != cudaSuccess) { rc = -2; break; }
        for (int s = 0; s < todo; ++s)
            if (cudaGraphLaunch(e->exec, e->stream) != cudaSuccess) { rc = -2; break; }
        if (rc) break;
        t_queued = now_ms();
        if (chain_host && cudaMemcpyAsync(chain_host + (size_t)done * per_step, E.chain, sizeof(double) * todo * per_step,
                                          cudaMemcpyDeviceToHost, e->stream) != cudaSuccess) rc = -2;
        if (lnprob_host && cudaMemcpyAsync(lnprob_host + (size_t)done * rows, E.chain_lnp, sizeof(double) * todo * rows,
                                           cudaMemcpyDeviceToHost, e->stream) != cudaSuccess) rc = -2;
        if (cudaStreamSynchronize(e->stream) != cudaSuccess) rc = -2;
        done += todo;
        e->steps_done += (unsigned int)todo;
    }
    E.chain = saved_chain;
    E.chain_lnp = saved_lnp;
    if (rc == 0 && n_accepted_host) {
        if (cudaMemcpy(n_accepted_host, E.n_accepted, sizeof(long long) * rows, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -2;
    }
    if (rc == 0 && cudaStreamSynchronize(e->stream) != cudaSuccess) rc = -2;
    if (rc == 0) rc = exchange_status(e->h, e->stream);
    if (timing)
        fprintf(stderr, "mcd_ensemble_run(%d steps, store %d): set-up %.2f ms | graph (re)build %.2f ms | %d graph launches queued "
                        "in %.2f ms | until done %.2f ms | total %.2f ms\n", n_steps, (int)store, t_graph - t_enter, t_loop - t_graph,
                n_steps, t_queued - t_loop, now_ms() - t_queued, now_ms() - t_enter);
    return rc;
}

extern "C" int mcd_ensemble_engine(const mcd_ensemble *e, int32_t *engine, int32_t *ctas_per_segment) {
    if (!e) return -1;
    if (engine) *engine = e->last_path;
    if (ctas_per_segment) *ctas_per_segment = e->last_path == 1 ? resident_chain_group(e->h) : 0;
    return 0;
}
extern "C" int mcd_ensemble_get_state(mcd_ensemble *e, double *pos_host, double *lnprob_host) {
    if (!e) return -1;
    ENS_CUDA(cudaSetDevice(e->device));
    ENS_CUDA(cudaStreamSynchronize(e->stream));
    const size_t rows = (size_t)e->E.n_walkers * e->E.n_segments;
    if (pos_host) ENS_CUDA(cudaMemcpy(pos_host, e->E.pos, sizeof(double) * rows * e->E.n_theta, cudaMemcpyDeviceToHost));
    if (lnprob_host) ENS_CUDA(cudaMemcpy(lnprob_host, e->E.lnp, sizeof(double) * rows, cudaMemcpyDeviceToHost));
    return 0;
}
