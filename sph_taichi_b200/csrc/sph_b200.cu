// sph_b200.cu -- host side of libsph_b200.so: context, workspace carving, launch sequences,
// CUDA-graph step replay, and the extern "C" ABI declared in include/sph_b200.h.
#include <cuda_runtime.h>
#ifndef SPH_EMU
#include <dlfcn.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "sph_dfsph.cuh"

namespace {

std::string g_create_error;

constexpr int NUM_TIMERS = 12;
constexpr int GRAPH_UNROLL = 4;  // must be even (ping-pong parity returns to its start)
const char *kTimerNames[NUM_TIMERS] = {"zero",  "hash",    "scan",  "bucket",    "rank_move", "boundary_volume",
                                       "density", "force", "advect_clamp", "rigid", "unused",    "total"};
enum { T_ZERO, T_HASH, T_SCAN, T_BUCKET, T_MOVE, T_BVOL, T_DENSITY, T_FORCE, T_ADVECT, T_RIGID, T_UNUSED, T_TOTAL };

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct Layout {
    uint64_t off_state[14];  // posm, veld, x0id, misc, acc (x2), aux, fpv (2 slots, contiguous), dfs
    uint64_t off_cid, off_grid_ids, off_perm, off_ticket;
    uint64_t off_zero_begin, off_tile_counter, off_tile_state, off_cell_end, off_zero_end;
    uint64_t off_solid_slot, off_status, off_bodies, off_scratch, off_nbr_list, off_nbr_cnt;
    int64_t npad;
    uint64_t total;
    int n_tiles;
};

Layout make_layout(int64_t n_max, int64_t C, int64_t n_solid, int n_bodies) {
    Layout L;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { uint64_t r = o; o = align_up(o + bytes, 256); return r; };
    uint64_t n = (uint64_t)(n_max > 0 ? n_max : 1);
    for (int k = 0; k < 14; ++k) {
        if (k == 12) { L.off_state[k] = L.off_state[11] + n * sizeof(float4); continue; }  // second half of fpv
        L.off_state[k] = take((k == 11 ? 2 : 1) * n * sizeof(float4));
    }
    L.off_cid = take(n * 4);
    L.off_grid_ids = take(n * 4);
    L.off_perm = take(n * 4);
    L.off_ticket = take(n * 4);
    L.n_tiles = (int)((C + 1 + SCAN_TILE - 1) / SCAN_TILE);  // +1: the slab-mode trash bucket
    // one contiguous region that a single memset clears every build
    L.off_zero_begin = o;
    L.off_tile_counter = take(256);
    L.off_tile_state = take((uint64_t)L.n_tiles * 8);
    L.off_cell_end = take((uint64_t)(C + 1) * 4);
    L.off_zero_end = o;
    L.off_solid_slot = take((uint64_t)(n_solid > 0 ? n_solid : 1) * 4);
    L.off_status = take(256);
    L.off_bodies = take((uint64_t)(n_bodies > 0 ? n_bodies : 1) * sizeof(RigidBodyDev));
    L.off_scratch = take(256);
    L.npad = (int64_t)align_up(n, 32);
    L.off_nbr_cnt = take((uint64_t)L.npad * 4);
    L.off_nbr_list = take((uint64_t)L.npad * 4 * NBR_CAP);
    L.total = o;
    return L;
}

#ifndef SPH_EMU
// NCCL point-to-point through dlopen: the library has no link-time dependency on NCCL (the CPU test-suite and
// single-GPU users never load it); inside a PyTorch process the already loaded libnccl.so.2 is reused.
struct NcclId { char internal[128]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
};
NcclApi g_nccl;
bool load_nccl() {
    if (g_nccl.lib) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_nccl.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) g_nccl.err = std::string("libnccl lacks ") + n; return p; };
    g_nccl.GetUniqueId = reinterpret_cast<int (*)(NcclId *)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<int (*)(void **, int, NcclId, int)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<int (*)(void *)>(sym("ncclCommDestroy"));
    g_nccl.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
    g_nccl.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
    g_nccl.Send = reinterpret_cast<int (*)(const void *, size_t, int, int, void *, cudaStream_t)>(sym("ncclSend"));
    g_nccl.Recv = reinterpret_cast<int (*)(void *, size_t, int, int, void *, cudaStream_t)>(sym("ncclRecv"));
    g_nccl.GetErrorString = reinterpret_cast<const char *(*)(int)>(sym("ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.GroupStart || !g_nccl.GroupEnd ||
        !g_nccl.Send || !g_nccl.Recv || !g_nccl.GetErrorString)
        return false;
    g_nccl.lib = h;
    return true;
}
#endif

}  // namespace

struct SphCtx {
    SphParams hp;
    DevParams P;
    DevArrays S;
    Layout L;
    int device = 0;
    int64_t n_max = 0, n_solid_cap = 0;
    int body_cap = 0;
    char *ws = nullptr;
    std::vector<SphRigidBody> bodies;
    bool has_dynamic_solids = true;  // conservative until pack() inspects the flags
    std::string err;
    int64_t launches = 0;
    // CUDA graphs, one per ping-pong parity
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    int64_t graph_kernels[2] = {0, 0};
    cudaGraphExec_t graph_multi[2] = {nullptr, nullptr};  // GRAPH_UNROLL steps per replay
    int64_t graph_multi_kernels[2] = {0, 0};
    int parity = 0;
    int var_density = 1, var_force = 1;  // 1 = production; 0 = ablation variants (SPH_DENSITY_VARIANT / SPH_FORCE_VARIANT)
    bool time_pair = false;
    cudaEvent_t pair_ev[3] = {nullptr, nullptr, nullptr};
    cudaStream_t capture_stream = nullptr;  // graphs are captured here (the legacy stream cannot capture)
    bool built = false;  // neighbour structure valid for current positions
    bool list_valid = false;  // neighbour lists valid for current positions (DFSPH kernels)
    // x-slab sharding (sph_shard_*)
    SphTransport transport{};
    bool has_transport = false;
    int rank = 0, world = 1;
    void *nccl_comm = nullptr;
    char *shard_buf = nullptr;  // library-owned: device step state + send staging (cudaMalloc at sph_shard_configure)
    bool shard_begun = false;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_packed = nullptr, ev_exchanged = nullptr;
    cudaGraphExec_t graph_shard[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [parity][wide exchange]
    int64_t graph_shard_kernels[2][2] = {{0, 0}, {0, 0}};
    int32_t narrow_cap = 0;      // records per side of the usual (sgw + 1 layers) exchange; halo_cap = sgw + 2 layers
    // hide the exchange behind the interior densities too (boundary densities first): pays off when a rank talks to
    // two neighbours and the exchange outlasts the interior force pass (8 ranks: 0.38 vs 0.2 ms); with one neighbour
    // or short exchanges the second density launch costs more than it hides (4 ranks: 0.457 vs 0.444 ms per step)
    bool split_density = false;
    int64_t shard_sequences = 0;  // sequences run so far (begin = 0): sequence q feeds the plan kernel of q + 1
};

namespace {

int fail(SphCtx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail((ctx), SPH_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

int validate_params(const SphParams *p, std::string &why) {
    if (!p) { why = "params is NULL"; return SPH_E_ARG; }
    if (p->dim != 3) { why = "only dim == 3 is supported (2-D is unreachable in the reference)"; return SPH_E_ARG; }
    for (int a = 0; a < 3; ++a)
        if (p->grid_num[a] < 3) { why = "grid_num must be >= 3 per axis"; return SPH_E_ARG; }
    if (!(p->h > 0.f) || !(p->density0 > 0.f)) { why = "h and density0 must be positive"; return SPH_E_ARG; }
    int64_t C = (int64_t)p->grid_num[0] * p->grid_num[1] * p->grid_num[2];
    if (C >= (1ll << 31)) { why = "too many grid cells for int32 indices"; return SPH_E_CAPACITY; }
    return SPH_OK;
}

void derive_params(SphCtx *c) { derive_dev_params(c->P, c->hp); }

void bind_arrays(SphCtx *c) {
    char *w = c->ws;
    const Layout &L = c->L;
    DevArrays &S = c->S;
    float4 *st[14];
    for (int k = 0; k < 14; ++k) st[k] = reinterpret_cast<float4 *>(w + L.off_state[k]);
    int p = c->parity;
    S.posm = st[0 + 5 * p]; S.veld = st[1 + 5 * p]; S.x0id = st[2 + 5 * p]; S.misc = st[3 + 5 * p]; S.acc = st[4 + 5 * p];
    int q = 1 - p;
    S.posm_n = st[0 + 5 * q]; S.veld_n = st[1 + 5 * q]; S.x0id_n = st[2 + 5 * q]; S.misc_n = st[3 + 5 * q]; S.acc_n = st[4 + 5 * q];
    S.aux = st[10];
    S.fpv = st[11];  // 2 float4 per particle
    S.dfs = st[13];
    S.cid = reinterpret_cast<int32_t *>(w + L.off_cid);
    S.grid_ids = reinterpret_cast<int32_t *>(w + L.off_grid_ids);
    S.perm = reinterpret_cast<int32_t *>(w + L.off_perm);
    S.tile_counter = reinterpret_cast<int32_t *>(w + L.off_tile_counter);
    S.tile_state = reinterpret_cast<unsigned long long *>(w + L.off_tile_state);
    S.cell_end = reinterpret_cast<int32_t *>(w + L.off_cell_end);
    S.ticket = reinterpret_cast<int32_t *>(w + L.off_ticket);
    S.solid_slot = reinterpret_cast<int32_t *>(w + L.off_solid_slot);
    S.status = reinterpret_cast<uint32_t *>(w + L.off_status);
    S.nbr_list = reinterpret_cast<int32_t *>(w + L.off_nbr_list);
    S.nbr_cnt = reinterpret_cast<int32_t *>(w + L.off_nbr_cnt);
    S.npad = (int32_t)L.npad;
    S.sd = nullptr;
    if (c->shard_buf) {  // {state ints | header x 2 | 2 sides x 4 arrays x halo_cap records}
        S.sd = reinterpret_cast<int32_t *>(c->shard_buf);
        char *q = c->shard_buf + 256;
        for (int side = 0; side < 2; ++side) { S.stage_hdr[side] = reinterpret_cast<int32_t *>(q); q += 256; }
        for (int side = 0; side < 2; ++side)
            for (int k = 0; k < 4; ++k) { S.stage[side][k] = reinterpret_cast<float4 *>(q); q += align_up((uint64_t)c->P.halo_cap * 16, 256); }
    }
}

inline RigidBodyDev *dev_bodies(SphCtx *c) { return reinterpret_cast<RigidBodyDev *>(c->ws + c->L.off_bodies); }

inline int blocks_for(int64_t n, int t) { return (int)((n + t - 1) / t); }
// Sharded steps: n is the capacity and the real counts are device state, so the LIGHT step kernels (histogram,
// bucket, rank + move) run grid-stride loops over the compact input numbering on a grid of `per_sm` resident blocks
// per SM instead of one thread per capacity slot (sort chain 0.121 -> 0.106 ms at 2 M particles per rank); the pair
// kernels keep one block per tile.
constexpr int kSMs = 148;
inline int step_grid(const SphCtx *c, int64_t n, int t, int per_sm) {
    const int full = blocks_for(n, t);
    return c->P.slab_on ? std::min(full, kSMs * per_sm) : full;
}

// Launch with programmatic stream serialisation (the kernels of the step chain start with pdl_wait()): the next
// kernel's blocks are scheduled while the previous grid drains, which removes most of the ~1.5 us gap per graph
// node.  SPH_PDL=0 switches it off (A/B).
#ifdef SPH_EMU
#define PDL_LAUNCH(kern, grid, block, st, ...) kern<<<grid, block, 0, st>>>(__VA_ARGS__)
#else
const bool g_pdl = !(std::getenv("SPH_PDL") && std::atoi(std::getenv("SPH_PDL")) == 0);
template <typename... KArgs, typename... Args>
inline void pdl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#define PDL_LAUNCH(kern, grid, block, st, ...) pdl_launch(kern, dim3(grid), dim3(block), st, __VA_ARGS__)
#endif

void drop_graphs(SphCtx *c) {
    for (int k = 0; k < 2; ++k) {
        if (c->graph[k]) { cudaGraphExecDestroy(c->graph[k]); c->graph[k] = nullptr; }
        if (c->graph_multi[k]) { cudaGraphExecDestroy(c->graph_multi[k]); c->graph_multi[k] = nullptr; }
        for (int w = 0; w < 2; ++w)
            if (c->graph_shard[k][w]) { cudaGraphExecDestroy(c->graph_shard[k][w]); c->graph_shard[k][w] = nullptr; }
    }
}

// ---- launch sequences -----------------------------------------------------------------
struct StageTimer {
    cudaEvent_t ev[NUM_TIMERS + 1];
    int stage_of[NUM_TIMERS + 1];
    int count = 0;
    bool on = false;
    cudaStream_t st = nullptr;
    void mark(int stage) {
        if (!on) return;
        cudaEventCreate(&ev[count]);
        cudaEventRecord(ev[count], st);
        stage_of[count] = stage;
        ++count;
    }
};

// particle_system.py:372-375.  Kernel count returned through *kernels.
int launch_neighbor_build(SphCtx *c, cudaStream_t st, StageTimer *tm, int64_t *kernels, bool move_acc = true) {
    const DevParams &P = c->P;
    if (P.n == 0) return SPH_OK;
    const Layout &L = c->L;
    if (tm) tm->mark(T_ZERO);
    CUDA_TRY(c, cudaMemsetAsync(c->ws + L.off_zero_begin, 0, L.off_zero_end - L.off_zero_begin, st));
    if (tm) tm->mark(T_HASH);
    PDL_LAUNCH(k_hash_count, step_grid(c, P.n, 256, 8), 256, st, P, c->S);
    if (tm) tm->mark(T_SCAN);
    PDL_LAUNCH(k_scan, L.n_tiles, SCAN_THREADS, st, c->S.cell_end, P.C + 1, c->S.tile_state, c->S.tile_counter);
    if (tm) tm->mark(T_BUCKET);
    PDL_LAUNCH(k_bucket, step_grid(c, P.n, 256, 8), 256, st, P, c->S);
    if (tm) tm->mark(T_MOVE);
    if (move_acc) PDL_LAUNCH(k_rank_move<true>, step_grid(c, P.n, 256, 8), 256, st, P, c->S);
    else PDL_LAUNCH(k_rank_move<false>, step_grid(c, P.n, 256, 8), 256, st, P, c->S);
    *kernels += 4;
    CUDA_TRY(c, cudaGetLastError());
    c->parity ^= 1;  // sorted state now lives in the other buffer set
    bind_arrays(c);
    c->built = true;
    return SPH_OK;
}

// density + neighbour lists, then forces (+ integration).  SPH_DENSITY_VARIANT / SPH_FORCE_VARIANT = 0
// select the ablation variants (second dense loop for the density sum; separate advect kernel).
void launch_pair_density(SphCtx *c, cudaStream_t st, int64_t *kernels, int split_mode = 0) {
    const DevParams &P = c->P;
    const int blocks = blocks_for(P.n, DENS_WARPS * 32);
    if (c->var_density == 0) PDL_LAUNCH((k_density_tma<false, false>), blocks, DENS_WARPS * 32, st, P, c->S, split_mode);
    else if (P.dfsph || c->var_density == 2) PDL_LAUNCH((k_density_tma<true, false>), blocks, DENS_WARPS * 32, st, P, c->S, split_mode);
    else PDL_LAUNCH((k_density_tma<true, true>), blocks, DENS_WARPS * 32, st, P, c->S, split_mode);
    *kernels += 1;
}
void launch_pair_force_and_advect(SphCtx *c, cudaStream_t st, StageTimer *tm, int64_t *kernels, int split_mode = 0) {
    const DevParams &P = c->P;
    if (P.uniform_fluid && c->var_force != 0) {
        if (P.slab_on)
            PDL_LAUNCH((k_force_packed<FORCE_BATCH, FORCE_THREADS, true, true>), blocks_for(P.n, FORCE_THREADS), FORCE_THREADS, st, P,
                       c->S, split_mode);
        else
            PDL_LAUNCH((k_force_packed<FORCE_BATCH, FORCE_THREADS, true>), blocks_for(P.n, FORCE_THREADS), FORCE_THREADS, st, P, c->S,
                       split_mode);
        *kernels += 1;
        if (tm) tm->mark(T_ADVECT);
        if (c->has_dynamic_solids && P.n_solid > 0) {
            k_advect_solids<<<blocks_for(P.n_solid, 256), 256, 0, st>>>(P, c->S);
            *kernels += 1;
        }
        return;
    }
    if (P.uniform_fluid) k_force_packed<FORCE_BATCH, FORCE_THREADS, false><<<blocks_for(P.n, FORCE_THREADS), FORCE_THREADS, 0, st>>>(P, c->S, 0);
    else k_force_general<4, 128><<<blocks_for(P.n, 128), 128, 0, st>>>(P, c->S);
    if (tm) tm->mark(T_ADVECT);
    k_advect<true><<<blocks_for(P.n, 256), 256, 0, st>>>(P, c->S);
    *kernels += 2;
}

int launch_boundary_volume(SphCtx *c, int moving, cudaStream_t st, int64_t *kernels) {
    const DevParams &P = c->P;
    if (P.n_solid == 0) return SPH_OK;
    k_boundary_volume<<<blocks_for((int64_t)P.n_solid * 32, 128), 128, 0, st>>>(P, c->S, moving);
    *kernels += 1;
    CUDA_TRY(c, cudaGetLastError());
    return SPH_OK;
}

int launch_rigid_solve(SphCtx *c, cudaStream_t st, int64_t *kernels) {
    // sph_base.py:247-260: per dynamic body solve_constraints, then enforce_boundary_3D(solid) -- all bodies in ONE
    // launch (one CTA per body carries the clamps of the whole loop for its own particles, see k_rigid), the dynamic
    // solids outside every body in a second one
    const int nb = (int)c->bodies.size();
    if (nb == 0) return SPH_OK;
    k_rigid<<<nb, RIGID_THREADS, 0, st>>>(c->P, c->S, dev_bodies(c), 0, 2, nullptr, nb);
    k_enforce_boundary_solid<<<blocks_for(c->P.n_solid, 256), 256, 0, st>>>(c->P, c->S, nb, dev_bodies(c), nb);
    *kernels += 2;
    CUDA_TRY(c, cudaGetLastError());
    return SPH_OK;
}

// One whole SPHBase.step() with the fused kernels (sph_base.py:263-271, WCSPH.py:152-156).
int launch_step(SphCtx *c, cudaStream_t st, StageTimer *tm, int64_t *kernels) {
    const DevParams &P = c->P;
    if (P.n == 0) return SPH_OK;
    int rc = launch_neighbor_build(c, st, tm, kernels, /*move_acc=*/false);
    if (rc) return rc;
    if (tm) tm->mark(T_BVOL);
    if (c->has_dynamic_solids) { rc = launch_boundary_volume(c, 1, st, kernels); if (rc) return rc; }
    if (tm) tm->mark(T_DENSITY);
    launch_pair_density(c, st, kernels);
    if (tm) tm->mark(T_FORCE);
    launch_pair_force_and_advect(c, st, tm, kernels);
    if (tm) tm->mark(T_RIGID);
    if (!c->bodies.empty()) { rc = launch_rigid_solve(c, st, kernels); if (rc) return rc; }
    if (tm) tm->mark(T_TOTAL);
    CUDA_TRY(c, cudaGetLastError());
    c->built = false; c->list_valid = false;  // positions moved
    return SPH_OK;
}

}  // namespace

// =========================================================================================
// extern "C" ABI
// =========================================================================================
extern "C" {

uint64_t sph_workspace_bytes(const SphParams *params, int64_t n_max, int64_t n_solid, int32_t n_bodies) {
    std::string why;
    if (validate_params(params, why) != SPH_OK || n_max < 0 || n_solid < 0 || n_bodies < 0) return 0;
    int64_t C = (int64_t)params->grid_num[0] * params->grid_num[1] * params->grid_num[2];
    return make_layout(n_max, C, n_solid, n_bodies).total;
}

int sph_create(const SphParams *params, int64_t n_max, int64_t n_solid, int32_t n_bodies, int32_t device,
               void *workspace, uint64_t workspace_bytes, SphCtx **out) {
    if (!out) return fail(nullptr, SPH_E_ARG, "out is NULL");
    *out = nullptr;
    std::string why;
    int rc = validate_params(params, why);
    if (rc) return fail(nullptr, rc, why);
    if (n_max < 0 || n_max >= (1ll << 31) - 1) return fail(nullptr, SPH_E_CAPACITY, "n_max out of int32 range");
    if ((uint64_t)align_up((uint64_t)(n_max > 0 ? n_max : 1), 32) * (uint64_t)NBR_CAP >= (1ull << 32))
        return fail(nullptr, SPH_E_CAPACITY, "n_max too large: neighbour-list slots are 32-bit (NBR_CAP * n_max < 2^32)");
    if (n_solid < 0 || n_bodies < 0) return fail(nullptr, SPH_E_ARG, "negative capacity");
    if (!workspace) return fail(nullptr, SPH_E_ARG, "workspace is NULL");
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(nullptr, SPH_E_ARG, "workspace must be 256-byte aligned");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SPH_E_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, SPH_E_ARG, "device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, SPH_E_CUDA, cudaGetErrorString(e));
    SphCtx *c = new (std::nothrow) SphCtx();
    if (!c) return fail(nullptr, SPH_E_CAPACITY, "out of host memory");
    c->hp = *params;
    c->device = device;
    c->n_max = n_max;
    c->n_solid_cap = n_solid;
    c->body_cap = n_bodies;
    int64_t C = (int64_t)params->grid_num[0] * params->grid_num[1] * params->grid_num[2];
    c->L = make_layout(n_max, C, n_solid, n_bodies);
    if (c->L.total > workspace_bytes) { delete c; return fail(nullptr, SPH_E_CAPACITY, "workspace too small"); }
    c->ws = static_cast<char *>(workspace);
    if (const char *v = std::getenv("SPH_DENSITY_VARIANT")) c->var_density = std::atoi(v);
    if (const char *v = std::getenv("SPH_FORCE_VARIANT")) c->var_force = std::atoi(v);
    c->P = DevParams{};
    if (const char *v = std::getenv("SPH_COLUMN_ORDER")) {  // experiments: a permutation of 012345678
        unsigned long long o = 0ull;
        unsigned seen = 0u;
        int k = 0;
        for (; v[k] >= '0' && v[k] <= '8' && k < 9; ++k) { o |= (unsigned long long)(v[k] - '0') << (4 * k); seen |= 1u << (v[k] - '0'); }
        if (k == 9 && seen == 0x1ffu && v[9] == 0) c->P.col_order = o;
        else { delete c; return fail(nullptr, SPH_E_ARG, "SPH_COLUMN_ORDER must be a permutation of 012345678"); }
    }
    derive_params(c);
    c->P.n = 0;
    c->P.n_solid = 0;
    bind_arrays(c);
    e = cudaMemset(c->ws + c->L.off_status, 0, 256);
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); delete c; return fail(nullptr, SPH_E_CUDA, m); }
    *out = c;
    return SPH_OK;
}

int sph_destroy(SphCtx *ctx) {
    if (!ctx) return SPH_OK;
    drop_graphs(ctx);
    if (ctx->capture_stream) cudaStreamDestroy(ctx->capture_stream);
#ifndef SPH_EMU
    if (ctx->nccl_comm && g_nccl.lib) g_nccl.CommDestroy(ctx->nccl_comm);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->ev_packed) cudaEventDestroy(ctx->ev_packed);
    if (ctx->ev_exchanged) cudaEventDestroy(ctx->ev_exchanged);
#endif
    if (ctx->shard_buf) cudaFree(ctx->shard_buf);
    delete ctx;
    return SPH_OK;
}

const char *sph_last_error(const SphCtx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sph_set_params(SphCtx *ctx, const SphParams *params) {
    if (!ctx) return SPH_E_ARG;
    std::string why;
    int rc = validate_params(params, why);
    if (rc) return fail(ctx, rc, why);
    for (int a = 0; a < 3; ++a)
        if (params->grid_num[a] != ctx->hp.grid_num[a]) return fail(ctx, SPH_E_ARG, "grid_num cannot change after creation");
    ctx->hp = *params;
    int n = ctx->P.n, ns = ctx->P.n_solid;
    derive_params(ctx);
    ctx->P.n = n; ctx->P.n_solid = ns;
    drop_graphs(ctx);
    return SPH_OK;
}

int sph_pack(SphCtx *ctx, const SphFields *f, int64_t n, void *stream) {
    if (!ctx || !f) return SPH_E_ARG;
    if (n < 0 || n > ctx->n_max) return fail(ctx, SPH_E_CAPACITY, "particle count exceeds n_max");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ctx->P.n = (int32_t)n;
    drop_graphs(ctx);
    ctx->built = false; ctx->list_valid = false;
    if (n == 0) return SPH_OK;
    k_pack<<<blocks_for(n, 256), 256, 0, st>>>(ctx->P, ctx->S, *f);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_set_solid_count(SphCtx *ctx, int64_t n_solid, int32_t has_dynamic_solids) {
    if (!ctx) return SPH_E_ARG;
    if (n_solid < 0 || n_solid > ctx->n_solid_cap) return fail(ctx, SPH_E_CAPACITY, "n_solid exceeds the capacity given at creation");
    ctx->P.n_solid = (int32_t)n_solid;
    ctx->has_dynamic_solids = has_dynamic_solids != 0;
    drop_graphs(ctx);
    return SPH_OK;
}

int sph_set_fluid_uniform(SphCtx *ctx, int32_t uniform, float fluid_m, float fluid_mV) {
    if (!ctx) return SPH_E_ARG;
    ctx->P.uniform_fluid = uniform ? 1 : 0;
    ctx->P.fluid_m = fluid_m;
    ctx->P.fluid_mV = fluid_mV;
    drop_graphs(ctx);
    return SPH_OK;
}

int sph_unpack(SphCtx *ctx, const SphFields *f, void *stream) {
    if (!ctx || !f) return SPH_E_ARG;
    if (ctx->P.n == 0) return SPH_OK;
    k_unpack<<<blocks_for(ctx->P.n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ctx->P, ctx->S, *f);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_unpack_xv(SphCtx *ctx, float *x, float *v, int32_t *object_id, void *stream) {
    if (!ctx || !x || !v) return SPH_E_ARG;
    if (ctx->P.n == 0) return SPH_OK;
    k_unpack_xv<<<blocks_for(ctx->P.n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ctx->P, ctx->S, x, v, object_id);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_upload_xv(SphCtx *ctx, const float *x, const float *v, void *stream) {
    if (!ctx || !x || !v) return SPH_E_ARG;
    if (ctx->P.n == 0) return SPH_OK;
    k_upload_xv<<<blocks_for(ctx->P.n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ctx->P, ctx->S, x, v);
    ctx->launches += 1;
    ctx->built = false; ctx->list_valid = false;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_copy_grid_particles_num(SphCtx *ctx, int32_t *out_dev, void *stream) {
    if (!ctx || !out_dev) return SPH_E_ARG;
    CUDA_TRY(ctx, cudaMemcpyAsync(out_dev, ctx->S.cell_end, (size_t)ctx->P.C * 4, cudaMemcpyDeviceToDevice,
                                  static_cast<cudaStream_t>(stream)));
    return SPH_OK;
}

int sph_neighbor_build(SphCtx *ctx, void *stream) {
    if (!ctx) return SPH_E_ARG;
    return launch_neighbor_build(ctx, static_cast<cudaStream_t>(stream), nullptr, &ctx->launches);
}

#define REQUIRE_BUILT(ctx)                                                                                  \
    if (!(ctx)) return SPH_E_ARG;                                                                           \
    if ((ctx)->P.n == 0) return SPH_OK;                                                                     \
    if (!(ctx)->built)                                                                                      \
        return fail((ctx), SPH_E_ARG, "neighbour structure is stale: call sph_neighbor_build first "        \
                                      "(the reference's kernels would read an outdated grid here)");

int sph_boundary_volume(SphCtx *ctx, int32_t moving, void *stream) {
    REQUIRE_BUILT(ctx);
    return launch_boundary_volume(ctx, moving, static_cast<cudaStream_t>(stream), &ctx->launches);
}

int sph_compute_densities(SphCtx *ctx, void *stream) {
    REQUIRE_BUILT(ctx);
    k_density_simple<false><<<blocks_for(ctx->P.n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(ctx->P, ctx->S);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_compute_non_pressure_forces(SphCtx *ctx, void *stream) {
    REQUIRE_BUILT(ctx);
    k_force_simple<true, false><<<blocks_for(ctx->P.n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(ctx->P, ctx->S);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_compute_pressure_forces(SphCtx *ctx, void *stream) {
    REQUIRE_BUILT(ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    k_eos<<<blocks_for(ctx->P.n, 256), 256, 0, st>>>(ctx->P, ctx->S);
    k_force_simple<false, true><<<blocks_for(ctx->P.n, 128), 128, 0, st>>>(ctx->P, ctx->S);
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_advect(SphCtx *ctx, void *stream) {
    if (!ctx) return SPH_E_ARG;
    if (ctx->P.n == 0) return SPH_OK;
    k_advect<false><<<blocks_for(ctx->P.n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ctx->P, ctx->S);
    ctx->launches += 1;
    ctx->built = false; ctx->list_valid = false;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_enforce_boundary(SphCtx *ctx, int32_t particle_type, void *stream) {
    if (!ctx) return SPH_E_ARG;
    if (ctx->P.n == 0) return SPH_OK;
    k_enforce_boundary<<<blocks_for(ctx->P.n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ctx->P, ctx->S, particle_type);
    ctx->launches += 1;
    ctx->built = false; ctx->list_valid = false;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_set_rigid_bodies(SphCtx *ctx, const SphRigidBody *bodies, int32_t n_bodies) {
    if (!ctx || n_bodies < 0 || (n_bodies > 0 && !bodies)) return SPH_E_ARG;
    if (n_bodies > ctx->body_cap) return fail(ctx, SPH_E_CAPACITY, "more rigid bodies than the workspace was sized for");
    for (int b = 0; b < n_bodies; ++b)  // validate everything before any state changes
        if (bodies[b].solid_begin < 0 || bodies[b].solid_end < bodies[b].solid_begin || bodies[b].solid_end > ctx->n_solid_cap)
            return fail(ctx, SPH_E_ARG, "rigid body solid-id range out of bounds");
    ctx->bodies.assign(bodies, bodies + n_bodies);
    std::vector<RigidBodyDev> h(n_bodies);
    for (int b = 0; b < n_bodies; ++b) {
        std::memset(&h[b], 0, sizeof(RigidBodyDev));
        h[b].object_id = bodies[b].object_id;
        h[b].solid_begin = bodies[b].solid_begin;
        h[b].solid_end = bodies[b].solid_end;
        for (int k = 0; k < 3; ++k) h[b].rest_cm[k] = bodies[b].rest_cm[k];
        h[b].R[0] = h[b].R[4] = h[b].R[8] = 1.0f;
    }
    if (n_bodies) CUDA_TRY(ctx, cudaMemcpy(dev_bodies(ctx), h.data(), sizeof(RigidBodyDev) * n_bodies, cudaMemcpyHostToDevice));
    drop_graphs(ctx);
    return SPH_OK;
}

static int rigid_call(SphCtx *ctx, int32_t body, int mode, float *out, void *stream) {
    if (!ctx) return SPH_E_ARG;
    if (body < 0 || body >= (int)ctx->bodies.size()) return fail(ctx, SPH_E_ARG, "rigid body index out of range");
    k_rigid<<<1, RIGID_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(ctx->P, ctx->S, dev_bodies(ctx), body, mode, out, 0);
    ctx->launches += 1;
    if (mode == 2) { ctx->built = false; ctx->list_valid = false; }  // modes 0 / 1 only read
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int sph_compute_com(SphCtx *ctx, int32_t body, float *out_dev, void *stream) {
    if (!out_dev) return SPH_E_ARG;
    return rigid_call(ctx, body, 0, out_dev, stream);
}
int sph_compute_rigid_rest_cm(SphCtx *ctx, int32_t body, void *stream) { return rigid_call(ctx, body, 1, nullptr, stream); }
int sph_solve_constraints(SphCtx *ctx, int32_t body, float *R_out_dev, void *stream) {
    return rigid_call(ctx, body, 2, R_out_dev, stream);
}

// Capture `nsteps` consecutive steps (starting from the current ping-pong parity) into one graph.
static int capture_steps(SphCtx *ctx, int nsteps, cudaGraphExec_t *out, int64_t *kernels_out) {
    const int par = ctx->parity;
    cudaGraph_t g = nullptr;
    int64_t kernels = 0;
    if (!ctx->capture_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->capture_stream, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->capture_stream;
    CUDA_TRY(ctx, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = SPH_OK;
    for (int s = 0; s < nsteps && rc == SPH_OK; ++s) rc = launch_step(ctx, cs, nullptr, &kernels);
    cudaError_t e = cudaStreamEndCapture(cs, &g);
    // capture advanced the host-side parity exactly as real steps do; rewind
    ctx->parity = par;
    bind_arrays(ctx);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(ctx, SPH_E_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(ctx, SPH_E_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
    *kernels_out = kernels;
    return SPH_OK;
}

int sph_get_rigid_state(SphCtx *ctx, int32_t body, float *out_dev12, void *stream) {
    if (!ctx || !out_dev12) return SPH_E_ARG;
    if (body < 0 || body >= (int)ctx->bodies.size()) return fail(ctx, SPH_E_ARG, "rigid body index out of range");
    const RigidBodyDev *B = dev_bodies(ctx) + body;
    CUDA_TRY(ctx, cudaMemcpyAsync(out_dev12, B->R, sizeof(float) * 12, cudaMemcpyDeviceToDevice,
                                  static_cast<cudaStream_t>(stream)));  // R[9] and cm[3] are adjacent
    return SPH_OK;
}

int sph_step(SphCtx *ctx, int32_t nsteps, void *stream) {
    if (!ctx || nsteps < 0) return SPH_E_ARG;
    if (ctx->P.dfsph) return fail(ctx, SPH_E_ARG, "sph_step is the fused WCSPH step; DFSPH is driven through sph_dfsph_op");
    if (ctx->P.n == 0) return SPH_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int s = 0;
    while (s < nsteps) {
        const int par = ctx->parity;
        // GRAPH_UNROLL (even) steps per replay amortise the graph-launch gap; the parity returns to `par`
        const bool multi = nsteps - s >= GRAPH_UNROLL;
        cudaGraphExec_t *slot = multi ? &ctx->graph_multi[par] : &ctx->graph[par];
        int64_t *kslot = multi ? &ctx->graph_multi_kernels[par] : &ctx->graph_kernels[par];
        const int span = multi ? GRAPH_UNROLL : 1;
        if (!*slot) {
            int rc = capture_steps(ctx, span, slot, kslot);
            if (rc) return rc;
        }
        CUDA_TRY(ctx, cudaGraphLaunch(*slot, st));
        ctx->launches += *kslot;
        if (span & 1) { ctx->parity = par ^ 1; bind_arrays(ctx); }
        ctx->built = false; ctx->list_valid = false;
        s += span;
    }
    return SPH_OK;
}

int sph_state_offsets(SphCtx *ctx, uint64_t *out5) {
    if (!ctx || !out5) return SPH_E_ARG;
    const float4 *cur[5] = {ctx->S.posm, ctx->S.veld, ctx->S.x0id, ctx->S.misc, ctx->S.acc};
    for (int k = 0; k < 5; ++k) out5[k] = (uint64_t)(reinterpret_cast<const char *>(cur[k]) - ctx->ws);
    return SPH_OK;
}

// =========================================================================================
// x-slab sharding (see include/sph_b200.h)
// =========================================================================================
}  // extern "C"
namespace {

#ifndef SPH_EMU
// SphTransport over the context's NCCL communicator (user = the SphCtx); ncclInt8 = 0
int nccl_group_start(void *) { return g_nccl.GroupStart(); }
int nccl_group_end(void *, void *) { return g_nccl.GroupEnd(); }
int nccl_send(void *user, const void *buf, uint64_t bytes, int32_t peer, void *stream) {
    return g_nccl.Send(buf, (size_t)bytes, 0, peer, static_cast<SphCtx *>(user)->nccl_comm, static_cast<cudaStream_t>(stream));
}
int nccl_recv(void *user, void *buf, uint64_t bytes, int32_t peer, void *stream) {
    return g_nccl.Recv(buf, (size_t)bytes, 0, peer, static_cast<SphCtx *>(user)->nccl_comm, static_cast<cudaStream_t>(stream));
}
#endif

// transport calls are host calls: under stream capture NCCL records its own kernels; the host-emulated build
// records the call itself so that a graph replay repeats it
#ifdef SPH_EMU
#define TRANSPORT_OP(st, expr) emu::submit((st), [=]() { (void)(expr); })
#else
#define TRANSPORT_OP(st, expr)                                                                                  \
    do {                                                                                                        \
        int _rc = (expr);                                                                                       \
        if (_rc) return fail(c, SPH_E_NCCL, std::string("halo exchange transport failed: ") + #expr + " -> " + std::to_string(_rc)); \
    } while (0)
#endif

// One halo exchange: the packed staging of both sides goes out, the neighbours' records land behind the live
// records of the CURRENT buffer set (left neighbour's at n - 2 * halo_cap, right neighbour's at n - halo_cap),
// their headers in the device step state.  Fixed-size messages (the record count travels in the header), so the
// whole group can be captured in a CUDA graph.
int shard_exchange(SphCtx *c, cudaStream_t st, bool wide) {
    if (c->world <= 1) return SPH_OK;
    if (!c->has_transport) return fail(c, SPH_E_ARG, "no transport: call sph_comm_init_nccl or sph_comm_set_transport first");
    const SphTransport t = c->transport;
    const DevArrays S = c->S;
    const uint64_t bytes = (uint64_t)(wide ? c->P.halo_cap : c->narrow_cap) * 16;
    float4 *cur[4] = {S.posm, S.veld, S.x0id, S.misc};
    TRANSPORT_OP(st, t.group_start(t.user));
    for (int side = 0; side < 2; ++side) {
        const int peer = side == 0 ? c->rank - 1 : c->rank + 1;
        if (peer < 0 || peer >= c->world) continue;
        const int64_t tail = (int64_t)c->P.n - (side == 0 ? 2 : 1) * (int64_t)c->P.halo_cap;
        for (int k = 0; k < 4; ++k) {
            const void *sp = S.stage[side][k];
            void *rp = cur[k] + tail;
            TRANSPORT_OP(st, t.send(t.user, sp, bytes, peer, st));
            TRANSPORT_OP(st, t.recv(t.user, rp, bytes, peer, st));
        }
        const void *hs = S.stage_hdr[side];
        void *hr = S.sd + (side == 0 ? SD_HDR_L : SD_HDR_R);
        TRANSPORT_OP(st, t.send(t.user, hs, 64, peer, st));
        TRANSPORT_OP(st, t.recv(t.user, hr, 64, peer, st));
    }
    TRANSPORT_OP(st, t.group_end(t.user, st));
    return SPH_OK;
}

// plan -> classify + sort -> info -> boundary densities -> boundary forces (-> staging) ->
// { exchange || interior densities -> apply -> interior forces }
// ev[8] (optional): CUDA events at the stage boundaries, [5..6] around the exchange, [7] after the interior
// densities (sph_shard_profile_step)
// wide: this sequence's exchange feeds a re-balancing step and carries sgw + 2 layers instead of sgw + 1
int shard_sequence(SphCtx *c, cudaStream_t st, bool compute, bool wide, int64_t *kernels, cudaEvent_t *ev = nullptr) {
    if (ev) cudaEventRecord(ev[0], st);
    k_shard_plan<<<1, 32, 0, st>>>(c->P, c->S);
    *kernels += 1;
    int rc = launch_neighbor_build(c, st, nullptr, kernels, /*move_acc=*/false);
    if (rc) return rc;
    k_shard_info<<<1, 32, 0, st>>>(c->P, c->S, c->P.sgw + (wide ? 2 : 1), wide ? c->P.halo_cap : c->narrow_cap);
    *kernels += 1;
    if (ev) cudaEventRecord(ev[1], st);
    const dim3 pack_grid(blocks_for(c->P.halo_cap, 256), 2);
    if (compute) {
        // boundary first: densities within one layer of the send ranges (split_density; otherwise all densities),
        // then forces + integration of the send ranges -- written into the send staging, not into the arrays: the
        // interior densities still read these positions
        launch_pair_density(c, st, kernels, /*split_mode=*/c->split_density ? 1 : 0);
        if (ev) cudaEventRecord(ev[2], st);
        launch_pair_force_and_advect(c, st, nullptr, kernels, /*split_mode=*/1);
    } else {
        if (ev) cudaEventRecord(ev[2], st);
        k_shard_apply<<<pack_grid, 256, 0, st>>>(c->P, c->S, /*copy_all=*/1);  // first exchange: the unchanged records
        *kernels += 1;
    }
    CUDA_TRY(c, cudaGetLastError());
    if (ev) cudaEventRecord(ev[3], st);
    // the exchange for the NEXT step runs on the communication stream while the interior is computed
#ifdef SPH_EMU
    cudaStream_t cs = st;  // the emulated runtime is synchronous
#else
    cudaStream_t cs = c->comm_stream;
    CUDA_TRY(c, cudaEventRecord(c->ev_packed, st));
    CUDA_TRY(c, cudaStreamWaitEvent(cs, c->ev_packed, 0));
#endif
    if (ev) cudaEventRecord(ev[5], cs);
    rc = shard_exchange(c, cs, wide);
    if (rc) return rc;
    if (ev) cudaEventRecord(ev[6], cs);
    if (compute) {
        if (c->split_density) launch_pair_density(c, st, kernels, /*split_mode=*/2);
        if (ev) cudaEventRecord(ev[7], st);
        k_shard_apply<<<pack_grid, 256, 0, st>>>(c->P, c->S, /*copy_all=*/0);  // boundary particles take their new state
        *kernels += 1;
        launch_pair_force_and_advect(c, st, nullptr, kernels, /*split_mode=*/2);
    }
#ifndef SPH_EMU
    CUDA_TRY(c, cudaEventRecord(c->ev_exchanged, cs));
    CUDA_TRY(c, cudaStreamWaitEvent(st, c->ev_exchanged, 0));
#endif
    if (ev) cudaEventRecord(ev[4], st);
    CUDA_TRY(c, cudaGetLastError());
    c->built = false; c->list_valid = false;
    return SPH_OK;
}

int capture_shard_step(SphCtx *ctx, bool wide, cudaGraphExec_t *out, int64_t *kernels_out) {
    const int par = ctx->parity;
    cudaGraph_t g = nullptr;
    int64_t kernels = 0;
    if (!ctx->capture_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->capture_stream, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->capture_stream;
    // relaxed: NCCL may issue CUDA calls of its own while its send / recv kernels are being captured
    CUDA_TRY(ctx, cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
    int rc = shard_sequence(ctx, cs, /*compute=*/true, wide, &kernels);
    cudaError_t e = cudaStreamEndCapture(cs, &g);
    ctx->parity = par;  // capture advanced the host-side parity exactly as a real step does; rewind
    bind_arrays(ctx);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(ctx, SPH_E_CUDA, std::string("sharded graph capture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(ctx, SPH_E_CUDA, std::string("sharded graph instantiate: ") + cudaGetErrorString(e));
    *kernels_out = kernels;
    return SPH_OK;
}

}  // namespace
extern "C" {

int sph_comm_unique_id(char out128[128]) {
#ifdef SPH_EMU
    (void)out128;
    return fail(nullptr, SPH_E_NCCL, "the host-emulated build has no NCCL: use sph_comm_set_transport");
#else
    if (!out128) return SPH_E_ARG;
    if (!load_nccl()) return fail(nullptr, SPH_E_NCCL, g_nccl.err);
    NcclId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc) return fail(nullptr, SPH_E_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
    std::memcpy(out128, id.internal, 128);
    return SPH_OK;
#endif
}

int sph_comm_init_nccl(SphCtx *ctx, const char id128[128], int32_t rank, int32_t world) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return SPH_E_ARG;
#ifdef SPH_EMU
    return fail(ctx, SPH_E_NCCL, "the host-emulated build has no NCCL: use sph_comm_set_transport");
#else
    if (!load_nccl()) return fail(ctx, SPH_E_NCCL, g_nccl.err);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    NcclId id;
    std::memcpy(id.internal, id128, 128);
    int rc = g_nccl.CommInitRank(&ctx->nccl_comm, world, id, rank);
    if (rc) return fail(ctx, SPH_E_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
    ctx->transport = SphTransport{ctx, nccl_group_start, nccl_group_end, nccl_send, nccl_recv};
    ctx->has_transport = true;
    ctx->rank = rank; ctx->world = world;
    drop_graphs(ctx);
    return SPH_OK;
#endif
}

int sph_comm_set_transport(SphCtx *ctx, const SphTransport *t, int32_t rank, int32_t world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return SPH_E_ARG;
    if (world > 1 && (!t || !t->group_start || !t->group_end || !t->send || !t->recv)) return fail(ctx, SPH_E_ARG, "incomplete transport");
    if (t) ctx->transport = *t;
    ctx->has_transport = t != nullptr;
    ctx->rank = rank; ctx->world = world;
    drop_graphs(ctx);
    return SPH_OK;
}

int sph_shard_configure(SphCtx *ctx, int32_t x_lo, int32_t x_hi, int32_t ghost_layers, int64_t halo_capacity,
                        int32_t rebalance_every) {
    if (!ctx) return SPH_E_ARG;
    if (x_lo < 0 || x_hi > ctx->P.gx || ghost_layers < 1 || x_hi - x_lo < ghost_layers + 2)
        return fail(ctx, SPH_E_ARG, "slab must lie inside the grid and be at least ghost_layers + 2 cell layers wide");
    if (!ctx->bodies.empty() || ctx->P.n_solid > 0)
        return fail(ctx, SPH_E_ARG, "x-slab sharding supports fluid-only scenes (rigid bodies are single-GPU, SURVEY 8e)");
    if (!ctx->P.uniform_fluid || ctx->P.dfsph)
        return fail(ctx, SPH_E_ARG, "x-slab sharding needs the uniform-fluid WCSPH step (one fluid density)");
    if (halo_capacity < 1 || 2 * halo_capacity + ctx->P.n > ctx->n_max)
        return fail(ctx, SPH_E_CAPACITY, "capacity: packed particles + 2 * halo_capacity receive slots exceed n_max");
    if (rebalance_every < 0) return SPH_E_ARG;
    const int32_t n_mine = ctx->P.n;
    drop_graphs(ctx);
    if (ctx->shard_buf) { cudaFree(ctx->shard_buf); ctx->shard_buf = nullptr; }
    const uint64_t bytes = 256 + 2 * 256 + 8 * align_up((uint64_t)halo_capacity * 16, 256);
    CUDA_TRY(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->shard_buf), bytes));
    CUDA_TRY(ctx, cudaMemset(ctx->shard_buf, 0, bytes));
    int32_t sd[SD_INTS] = {0};
    sd[SD_N_LIVE] = n_mine; sd[SD_OWNED] = n_mine; sd[SD_SX0] = x_lo; sd[SD_SX1] = x_hi;
    CUDA_TRY(ctx, cudaMemcpy(ctx->shard_buf, sd, sizeof(sd), cudaMemcpyHostToDevice));
    ctx->P.slab_on = 1; ctx->P.sgw = ghost_layers; ctx->P.halo_cap = (int32_t)halo_capacity;
    ctx->P.has_left = ctx->rank > 0; ctx->P.has_right = ctx->rank + 1 < ctx->world;
    ctx->P.rebalance_every = rebalance_every;
    ctx->narrow_cap = (int32_t)((halo_capacity * (ghost_layers + 1) + ghost_layers + 1) / (ghost_layers + 2));
    ctx->shard_sequences = 0;
    ctx->split_density = ctx->world > 4;
    if (const char *v = std::getenv("SPH_SHARD_SPLIT_DENSITY")) ctx->split_density = std::atoi(v) != 0;
    ctx->P.n = (int32_t)ctx->n_max;  // from here on n is the CAPACITY; the live count is device state
    ctx->shard_begun = false;
    bind_arrays(ctx);
#ifndef SPH_EMU
    if (!ctx->comm_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // the exchange must get SMs while the interior pass runs
        CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_packed, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_exchanged, cudaEventDisableTiming));
    }
#endif
    return SPH_OK;
}

// the exchange at the end of sequence q delivers the records plan kernel q + 1 works on; that one re-balances
static inline bool shard_wide(const SphCtx *c) {
    const int64_t next_plan = c->shard_sequences + 1;
    return c->P.rebalance_every > 0 && next_plan % c->P.rebalance_every == 0;
}

#define REQUIRE_SHARD(ctx)                                                                              \
    if (!(ctx)) return SPH_E_ARG;                                                                       \
    if (!(ctx)->P.slab_on || !(ctx)->shard_buf) return fail((ctx), SPH_E_ARG, "sph_shard_configure was not called");

int sph_shard_begin(SphCtx *ctx, void *stream) {
    REQUIRE_SHARD(ctx);
    ctx->shard_sequences = 0;
    int rc = shard_sequence(ctx, static_cast<cudaStream_t>(stream), /*compute=*/false, shard_wide(ctx), &ctx->launches);
    if (rc) return rc;
    ctx->shard_sequences = 1;
    ctx->shard_begun = true;
    return SPH_OK;
}

int sph_halo_exchange(SphCtx *ctx, void *stream) {
    REQUIRE_SHARD(ctx);
    return shard_exchange(ctx, static_cast<cudaStream_t>(stream), /*wide=*/true);
}

int sph_shard_step(SphCtx *ctx, int32_t nsteps, void *stream) {
    REQUIRE_SHARD(ctx);
    if (nsteps < 0) return SPH_E_ARG;
    if (!ctx->shard_begun) return fail(ctx, SPH_E_ARG, "sph_shard_begin was not called");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Un-graphed launches are the DEFAULT here: nothing in a sharded step waits for the host any more, so the launches
    // run ahead of the device, and a graph replay -- whose two branches (exchange || interior forces) must both finish
    // before the next replay may start -- measured 9-12 % SLOWER (0.779 vs 0.712 ms per step at 2 x 2 M particles,
    // 0.51 vs 0.45 at 4 x 1 M; profiles/r02_shard_timing.txt).  SPH_SHARD_GRAPH=1 replays the captured graph instead
    // (one graph per ping-pong parity and exchange width; pays off for small per-rank problems).
    static const bool eager = !(std::getenv("SPH_SHARD_GRAPH") && std::atoi(std::getenv("SPH_SHARD_GRAPH")) != 0);
    for (int s = 0; s < nsteps; ++s) {
        const bool wide = shard_wide(ctx);
        if (eager) {
            int rc = shard_sequence(ctx, st, /*compute=*/true, wide, &ctx->launches);
            if (rc) return rc;
            ctx->shard_sequences += 1;
            continue;
        }
        const int par = ctx->parity;
        if (!ctx->graph_shard[par][wide]) {
            int rc = capture_shard_step(ctx, wide, &ctx->graph_shard[par][wide], &ctx->graph_shard_kernels[par][wide]);
            if (rc) return rc;
        }
        CUDA_TRY(ctx, cudaGraphLaunch(ctx->graph_shard[par][wide], st));
        ctx->launches += ctx->graph_shard_kernels[par][wide];
        ctx->parity = par ^ 1;
        bind_arrays(ctx);
        ctx->shard_sequences += 1;
    }
    return SPH_OK;
}

int sph_shard_info(SphCtx *ctx, int32_t *out16, uint64_t *out_sent, void *stream) {
    REQUIRE_SHARD(ctx);
    if (!out16) return SPH_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t sd[SD_INTS];
    uint32_t status = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(sd, ctx->S.sd, sizeof(sd), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(&status, ctx->S.status, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    for (int k = 0; k < 14; ++k) out16[k] = sd[k];
    out16[14] = (int32_t)status; out16[15] = 0;
    if (out_sent) *out_sent = ((uint64_t)(uint32_t)sd[SD_SENT_HI] << 32) | (uint32_t)sd[SD_SENT_LO];
    return SPH_OK;
}

int sph_shard_profile_step(SphCtx *ctx, float *ms_out6, void *stream) {
    REQUIRE_SHARD(ctx);
    if (!ms_out6) return SPH_E_ARG;
    if (!ctx->shard_begun) return fail(ctx, SPH_E_ARG, "sph_shard_begin was not called");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaEvent_t ev[8];
    for (int k = 0; k < 8; ++k) CUDA_TRY(ctx, cudaEventCreate(&ev[k]));
    int rc = shard_sequence(ctx, st, /*compute=*/true, shard_wide(ctx), &ctx->launches, ev);
    if (rc) return rc;
    ctx->shard_sequences += 1;
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ms_out6[0], ev[0], ev[1]);  // plan + sort + info
    cudaEventElapsedTime(&ms_out6[1], ev[1], ev[2]);  // boundary densities
    cudaEventElapsedTime(&ms_out6[2], ev[2], ev[3]);  // boundary forces (-> staging)
    cudaEventElapsedTime(&ms_out6[3], ev[3], ev[7]);  // interior densities
    cudaEventElapsedTime(&ms_out6[4], ev[7], ev[4]);  // apply + interior forces + waiting for the exchange, if any
    ms_out6[5] = 0.f;
    if (ctx->world > 1) cudaEventElapsedTime(&ms_out6[5], ev[5], ev[6]);
    for (int k = 0; k < 8; ++k) cudaEventDestroy(ev[k]);
    return SPH_OK;
}

int sph_set_dfsph(SphCtx *ctx, int32_t enable) {
    if (!ctx) return SPH_E_ARG;
    ctx->P.dfsph = enable ? 1 : 0;
    drop_graphs(ctx);
    return SPH_OK;
}

int sph_dfsph_op(SphCtx *ctx, int32_t op, float arg, void *out_dev, void *stream) {
    if (!ctx) return SPH_E_ARG;
    if (!ctx->P.dfsph) return fail(ctx, SPH_E_ARG, "sph_set_dfsph(1) was not called");
    const DevParams &P = ctx->P;
    if (P.n == 0) return SPH_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int b128 = blocks_for(P.n, 128), b256 = blocks_for(P.n, 256);
    if (op >= DFSPH_COMPUTE_DENSITIES && op <= DFSPH_NON_PRESSURE_FORCES && op != DFSPH_DENSITY_ERROR &&
        op != DFSPH_MULTIPLY_FACTOR) {
        if (!ctx->built)
            return fail(ctx, SPH_E_ARG, "neighbour structure is stale: call sph_neighbor_build first");
        if (op != DFSPH_COMPUTE_DENSITIES && !ctx->list_valid)
            return fail(ctx, SPH_E_ARG, "neighbour lists are stale: run op 0 (compute_densities) first");
    }
    switch (op) {
        case DFSPH_COMPUTE_DENSITIES:
            launch_pair_density(ctx, st, &ctx->launches);
            ctx->list_valid = true;
            break;
        case DFSPH_COMPUTE_FACTOR: k_dfsph_factor<<<b128, 128, 0, st>>>(P, ctx->S); break;
        case DFSPH_DENSITY_CHANGE: k_dfsph_density_change<0><<<b128, 128, 0, st>>>(P, ctx->S); break;
        case DFSPH_DENSITY_ADV: k_dfsph_density_change<1><<<b128, 128, 0, st>>>(P, ctx->S); break;
        case DFSPH_DENSITY_ERROR:
            if (!out_dev) return SPH_E_ARG;
            k_dfsph_density_error<<<b256, 256, 0, st>>>(P, ctx->S, arg, static_cast<double *>(out_dev));
            break;
        case DFSPH_MULTIPLY_FACTOR: k_dfsph_multiply_factor<<<b256, 256, 0, st>>>(P, ctx->S, arg); break;
        case DFSPH_DIVERGENCE_ITERATION: k_dfsph_iteration<0><<<b128, 128, 0, st>>>(P, ctx->S, nullptr); break;
        case DFSPH_PRESSURE_ITERATION: k_dfsph_iteration<1><<<b128, 128, 0, st>>>(P, ctx->S, nullptr); break;
        case DFSPH_NON_PRESSURE_FORCES: k_dfsph_non_pressure<<<b128, 128, 0, st>>>(P, ctx->S); break;
        case DFSPH_PREDICT_VELOCITY: k_dfsph_predict_velocity<<<b256, 256, 0, st>>>(P, ctx->S); break;
        case DFSPH_ADVECT:
            k_dfsph_advect<<<b256, 256, 0, st>>>(P, ctx->S);
            ctx->built = false; ctx->list_valid = false;
            break;
        default: return fail(ctx, SPH_E_ARG, "unknown DFSPH op");
    }
    if (op != DFSPH_COMPUTE_DENSITIES) ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

}  // extern "C"
namespace {

// The Jacobi loop of divergence_solve (mode 0, DFSPH.py:245-254) or pressure_solve (mode 1, DFSPH.py:323-331) with the
// loop condition on the device: sweeps are launched in batches (the first as long as the caller expects the loop to
// run -- the previous step's count is a good guess), every sweep ends with k_dfsph_check, sweeps behind the
// converged one return at once, and the host reads the control block once per batch.
int dfsph_solve_loop(SphCtx *ctx, int mode, int32_t max_iterations, double eta, float offset, int64_t n_fluid,
                     int32_t first_batch, DfsphCtrl *h, cudaStream_t st) {
    const DevParams &P = ctx->P;
    DfsphCtrl *ctrl = reinterpret_cast<DfsphCtrl *>(ctx->ws + ctx->L.off_scratch);
    *h = DfsphCtrl{};
    CUDA_TRY(ctx, cudaMemsetAsync(ctrl, 0, sizeof(DfsphCtrl), st));
    if (P.n == 0) return SPH_OK;
    const int b128 = blocks_for(P.n, 128);
    int batch = std::max(1, std::min(first_batch, 64));
    while (!h->done) {
        for (int s = 0; s < batch; ++s) {
            if (mode == 0) {
                k_dfsph_iteration<0><<<b128, 128, 0, st>>>(P, ctx->S, ctrl);
                k_dfsph_density_change_err<0><<<b128, 128, 0, st>>>(P, ctx->S, offset, ctrl);
            } else {
                k_dfsph_iteration<1><<<b128, 128, 0, st>>>(P, ctx->S, ctrl);
                k_dfsph_density_change_err<1><<<b128, 128, 0, st>>>(P, ctx->S, offset, ctrl);
            }
            k_dfsph_check<<<1, 1, 0, st>>>(ctrl, (double)n_fluid, eta, max_iterations);
        }
        ctx->launches += 3 * (int64_t)batch;
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaMemcpyAsync(h, ctrl, sizeof(DfsphCtrl), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        batch = 2;  // the guess was short: continue in pairs
    }
    return SPH_OK;
}

}  // namespace
extern "C" {

int sph_dfsph_solve(SphCtx *ctx, int32_t mode, int32_t max_iterations, double eta, float offset, int64_t n_fluid,
                    int32_t first_batch, int32_t *iterations_out, int32_t *sweeps_out, double *avg_err_out, void *stream) {
    if (!ctx || (mode != 0 && mode != 1) || n_fluid < 1) return SPH_E_ARG;
    if (!ctx->P.dfsph) return fail(ctx, SPH_E_ARG, "sph_set_dfsph(1) was not called");
    if (!ctx->built || !ctx->list_valid)
        return fail(ctx, SPH_E_ARG, "neighbour lists are stale: run sph_neighbor_build and op 0 (compute_densities) first");
    DfsphCtrl h{};
    int rc = dfsph_solve_loop(ctx, mode, max_iterations, eta, offset, n_fluid, first_batch, &h, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    if (iterations_out) *iterations_out = h.iterations;
    if (sweeps_out) *sweeps_out = h.sweeps;
    if (avg_err_out) *avg_err_out = h.last_avg;
    return SPH_OK;
}

int sph_dfsph_step(SphCtx *ctx, int32_t nsteps, SphDfsphStep *io, void *stream) {
    if (!ctx || !io || nsteps < 0 || io->n_fluid < 1) return SPH_E_ARG;
    if (!ctx->P.dfsph) return fail(ctx, SPH_E_ARG, "sph_set_dfsph(1) was not called");
    if (ctx->P.slab_on) return fail(ctx, SPH_E_ARG, "DFSPH is single-GPU");
    if (ctx->P.n == 0) return SPH_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int s = 0; s < nsteps; ++s) {
        // ps.initialize_particle_system(), compute_moving_boundary_volume() (sph_base.py:264-265)
        int rc = launch_neighbor_build(ctx, st, nullptr, &ctx->launches);
        if (rc) return rc;
        if (ctx->has_dynamic_solids) { rc = launch_boundary_volume(ctx, 1, st, &ctx->launches); if (rc) return rc; }
        const DevParams &P = ctx->P;  // after the build: the array binding changed sides
        const int b128 = blocks_for(P.n, 128), b256 = blocks_for(P.n, 256);
        // substep (DFSPH.py:399-408)
        launch_pair_density(ctx, st, &ctx->launches);
        ctx->list_valid = true;
        k_dfsph_factor<<<b128, 128, 0, st>>>(P, ctx->S);
        ctx->launches += 1;
        DfsphCtrl h{};
        if (io->enable_divergence_solver) {  // divergence_solve, DFSPH.py:236-276
            k_dfsph_density_change<0><<<b128, 128, 0, st>>>(P, ctx->S);
            k_dfsph_multiply_factor<<<b256, 256, 0, st>>>(P, ctx->S, io->inv_dt);
            ctx->launches += 2;
            rc = dfsph_solve_loop(ctx, 0, io->max_iterations_v, io->eta_v, 0.0f, io->n_fluid, io->first_batch_v, &h, st);
            if (rc) return rc;
            io->iterations_v = h.iterations; io->avg_err_v = h.last_avg; io->first_batch_v = std::max(1, h.sweeps);
            k_dfsph_multiply_factor<<<b256, 256, 0, st>>>(P, ctx->S, io->dt);
            ctx->launches += 1;
        }
        k_dfsph_non_pressure<<<b128, 128, 0, st>>>(P, ctx->S);
        k_dfsph_predict_velocity<<<b256, 256, 0, st>>>(P, ctx->S);
        // pressure_solve, DFSPH.py:314-352
        k_dfsph_density_change<1><<<b128, 128, 0, st>>>(P, ctx->S);
        k_dfsph_multiply_factor<<<b256, 256, 0, st>>>(P, ctx->S, io->inv_dt2);
        ctx->launches += 4;
        rc = dfsph_solve_loop(ctx, 1, io->max_iterations, io->eta, io->density0, io->n_fluid, io->first_batch, &h, st);
        if (rc) return rc;
        io->iterations = h.iterations; io->avg_err = h.last_avg; io->first_batch = std::max(1, h.sweeps);
        k_dfsph_advect<<<b256, 256, 0, st>>>(P, ctx->S);
        ctx->launches += 1;
        ctx->built = false; ctx->list_valid = false;
        // solve_rigid_body(), enforce_boundary_3D(fluid) (sph_base.py:270-271)
        if (!ctx->bodies.empty()) { rc = launch_rigid_solve(ctx, st, &ctx->launches); if (rc) return rc; }
        k_enforce_boundary<<<b256, 256, 0, st>>>(P, ctx->S, /*particle_type=*/1);
        ctx->launches += 1;
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return SPH_OK;
}

int sph_read_status(SphCtx *ctx, uint32_t *status_out, void *stream) {
    if (!ctx || !status_out) return SPH_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(status_out, ctx->S.status, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return SPH_OK;
}

int sph_clear_status(SphCtx *ctx, void *stream) {
    if (!ctx) return SPH_E_ARG;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->S.status, 0, 4, static_cast<cudaStream_t>(stream)));
    return SPH_OK;
}

int sph_neighbor_stats(SphCtx *ctx, int32_t *out_dev4, void *stream) {
    if (!ctx || !out_dev4) return SPH_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(ctx, cudaMemsetAsync(out_dev4, 0, 16, st));
    if (ctx->P.n == 0) return SPH_OK;
    k_neighbor_stats<<<blocks_for(ctx->P.n, 256), 256, 0, st>>>(ctx->P, ctx->S, out_dev4);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return SPH_OK;
}

int64_t sph_particle_count(const SphCtx *ctx) { return ctx ? ctx->P.n : 0; }
int64_t sph_launch_count(const SphCtx *ctx) { return ctx ? ctx->launches : 0; }

int sph_profile_step(SphCtx *ctx, float *ms, int32_t n, void *stream) {
    if (!ctx || !ms || n <= 0) return SPH_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StageTimer tm;
    tm.on = true;
    tm.st = st;
    int rc = launch_step(ctx, st, &tm, &ctx->launches);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    float out[NUM_TIMERS];
    for (int k = 0; k < NUM_TIMERS; ++k) out[k] = 0.f;
    for (int k = 0; k + 1 < tm.count; ++k) {
        float t = 0.f;
        cudaEventElapsedTime(&t, tm.ev[k], tm.ev[k + 1]);
        out[tm.stage_of[k]] += t;
    }
    if (tm.count >= 2) cudaEventElapsedTime(&out[T_TOTAL], tm.ev[0], tm.ev[tm.count - 1]);
    for (int k = 0; k < tm.count; ++k) cudaEventDestroy(tm.ev[k]);
    int m = n < NUM_TIMERS ? n : NUM_TIMERS;
    for (int k = 0; k < m; ++k) ms[k] = out[k];
    return m;
}

const char *sph_timer_name(int32_t i) { return (i >= 0 && i < NUM_TIMERS) ? kTimerNames[i] : ""; }

}  // extern "C"
