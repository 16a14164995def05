// group.cu — trajectory sharding across the GPUs of one box behind the C ABI (SURVEY.md §8(b) "Multi-GPU", §8(e)).
//
// Trajectories are independent (nothing in rk_step, handle_step_adaptive, cfm_general or magnus_42 couples them), so
// stepping needs no collective: rank r integrates the contiguous range shard_range(N, r, G) of the ensemble. NCCL (over
// NVLink 5 / NVSwitch) appears only at the two ends of a solve:
//   vo_group_gather        final states of every shard -> one device ensemble on the root -> ONE device-to-host copy
//   vo_group_reduce_stats  per-rank counters -> all-reduce (sums of accepted / rejected / status flags, min / max of t)
// plus vo_group_allreduce for the one scalar a domain-decomposed adaptive step shares (the partial error norm).
//
// Two ways to form a group, same entry points afterwards:
//   * one process per GPU (the launch model of bench.py and torchrun): rank 0 calls vo_group_unique_id, the host program
//     ships the 128 bytes to the other ranks by whatever means it has, every rank calls vo_group_create_rank;
//   * one process, one host thread, several GPUs: vo_group_create_local over one ctx per device (ncclCommInitAll); the
//     per-member arguments are then arrays with one entry per GPU and every collective is issued inside
//     ncclGroupStart / ncclGroupEnd.
// libnccl is dlopen'ed on first use (the copy already mapped into the process, e.g. torch's, else the system one), so the
// library loads and single-GPU work runs without it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "rk_small.cuh"

// solver.cu
int32_t vo_solver_local_stats(vo_solver s, unsigned long long* sums_dev /*[6]*/, double* mm_dev /*[2]*/);

namespace {

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*getUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*commInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*commInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*commDestroy)(ncclComm_t) = nullptr;
    const char* (*getErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*groupStart)() = nullptr;
    ncclResult_t (*groupEnd)() = nullptr;
    ncclResult_t (*send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*allReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*allGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*getVersion)(int*) = nullptr;
};

bool nccl_load(Nccl** out, std::string& err) {
    static Nccl n;
    static std::mutex mu;
    static bool tried = false, ok = false;
    static std::string why;
    std::lock_guard<std::mutex> lock(mu);
    if (!tried) {
        tried = true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            n.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.h) break;
        }
        if (!n.h) {
            const char* de = dlerror();  // one call: dlerror() clears the message it returns
            why = std::string("NCCL is not available (dlopen libnccl.so.2: ") + (de ? de : "?") + ")";
        } else {
            ok = true;
#define VO_NCCL_SYM(field, sym)                                  \
    n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.h, sym)); \
    if (!n.field) ok = false, why = std::string("NCCL: missing symbol ") + sym;
            VO_NCCL_SYM(getUniqueId, "ncclGetUniqueId")
            VO_NCCL_SYM(commInitRank, "ncclCommInitRank")
            VO_NCCL_SYM(commInitAll, "ncclCommInitAll")
            VO_NCCL_SYM(commDestroy, "ncclCommDestroy")
            VO_NCCL_SYM(getErrorString, "ncclGetErrorString")
            VO_NCCL_SYM(groupStart, "ncclGroupStart")
            VO_NCCL_SYM(groupEnd, "ncclGroupEnd")
            VO_NCCL_SYM(send, "ncclSend")
            VO_NCCL_SYM(recv, "ncclRecv")
            VO_NCCL_SYM(allReduce, "ncclAllReduce")
            VO_NCCL_SYM(allGather, "ncclAllGather")
            VO_NCCL_SYM(broadcast, "ncclBroadcast")
            VO_NCCL_SYM(getVersion, "ncclGetVersion")
#undef VO_NCCL_SYM
        }
    }
    err = why;
    *out = &n;
    return ok;
}

struct Member {
    vo_ctx ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0;
    double* gbuf = nullptr;  // root: the gathered ensemble [d][n_total]
    size_t gbuf_elems = 0;
    void* scratch = nullptr;  // 256 bytes: counters for reduce_stats / allreduce
    void* pinned = nullptr;   // 256 bytes
};

}  // namespace

struct vo_group_s {
    Nccl* nc = nullptr;
    int world = 1;
    std::vector<Member> m;  // local members: 1 in the process-per-GPU model, `world` in the single-process model
    std::string err;
};

namespace {

int32_t g_fail(vo_group g, int32_t code, const std::string& msg) {
    if (g) g->err = msg;
    for (size_t i = 0; g && i < g->m.size(); ++i)
        if (g->m[i].ctx) g->m[i].ctx->err = msg;
    g_vo_tls_err = msg;
    return code;
}

#define VO_NCCL(g, call)                                                                                                 \
    do {                                                                                                                  \
        ncclResult_t _r = (call);                                                                                         \
        if (_r != ncclSuccess) return g_fail((g), VO_ERR_NCCL, std::string(#call) + ": " + (g)->nc->getErrorString(_r)); \
    } while (0)

// ncclGroupStart / ncclGroupEnd as a scope: an error return between the two must not leave the NCCL group open (every later NCCL call
// of the process would be queued into it and never issued).
struct GroupScope {
    Nccl* nc = nullptr;
    bool open = false;
    ncclResult_t start(Nccl* n) {
        nc = n;
        const ncclResult_t r = nc->groupStart();
        open = r == ncclSuccess;
        return r;
    }
    ncclResult_t end() {
        open = false;
        return nc->groupEnd();
    }
    ~GroupScope() {
        if (open) nc->groupEnd();
    }
};

int32_t member_alloc(vo_group g) {
    for (Member& mb : g->m) {
        DeviceGuard dg(mb.ctx->device);
        if (vo_dmalloc(&mb.scratch, 256) != cudaSuccess || cudaMallocHost(&mb.pinned, 256) != cudaSuccess) return g_fail(g, VO_ERR_ALLOC, "vo_group: scratch allocation failed");
    }
    return VO_OK;
}

void shard_range(int64_t n_total, int rank, int world, int64_t* lo, int64_t* hi) {
    const int64_t per = (n_total + world - 1) / world;  // contiguous ceil(N/G) ranges
    *lo = std::min<int64_t>(n_total, (int64_t)rank * per);
    *hi = std::min<int64_t>(n_total, *lo + per);
}

Member* find_member(vo_group g, int rank) {
    for (Member& mb : g->m)
        if (mb.rank == rank) return &mb;
    return nullptr;
}

}  // namespace

extern "C" {

int32_t vo_group_unique_id(void* id_out) {
    if (!id_out) return vo_fail(nullptr, VO_ERR_BAD_ARG, "vo_group_unique_id: NULL argument");
    Nccl* nc = nullptr;
    std::string err;
    if (!nccl_load(&nc, err)) return vo_fail(nullptr, VO_ERR_UNSUPPORTED, err);
    static_assert(sizeof(ncclUniqueId) == VO_GROUP_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = nc->getUniqueId(&id);
    if (r != ncclSuccess) return vo_fail(nullptr, VO_ERR_NCCL, std::string("ncclGetUniqueId: ") + nc->getErrorString(r));
    std::memcpy(id_out, &id, sizeof id);
    return VO_OK;
}

int32_t vo_group_create_rank(vo_ctx ctx, const void* id, int32_t rank, int32_t world, vo_group* out) {
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) return vo_fail(ctx, VO_ERR_BAD_ARG, "vo_group_create_rank: bad argument");
    vo_group g = new vo_group_s();
    g->world = world;
    g->m.resize(1);
    g->m[0].ctx = ctx, g->m[0].rank = rank;
    if (world > 1) {
        std::string err;
        if (!nccl_load(&g->nc, err)) {
            delete g;
            return vo_fail(ctx, VO_ERR_UNSUPPORTED, err);
        }
        DeviceGuard dg(ctx->device);
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof uid);
        ncclResult_t r = g->nc->commInitRank(&g->m[0].comm, world, uid, rank);
        if (r != ncclSuccess) {
            const std::string msg = std::string("ncclCommInitRank: ") + g->nc->getErrorString(r);
            delete g;
            return vo_fail(ctx, VO_ERR_NCCL, msg);
        }
    }
    int32_t rc = member_alloc(g);
    if (rc != VO_OK) {
        vo_group_destroy(g);
        return rc;
    }
    *out = g;
    return VO_OK;
}

int32_t vo_group_create_local(const vo_ctx* ctxs, int32_t n, vo_group* out) {
    if (!ctxs || !out || n < 1) return vo_fail(nullptr, VO_ERR_BAD_ARG, "vo_group_create_local: bad argument");
    for (int i = 0; i < n; ++i)
        if (!ctxs[i]) return vo_fail(nullptr, VO_ERR_BAD_ARG, "vo_group_create_local: NULL ctx");
    vo_group g = new vo_group_s();
    g->world = n;
    g->m.resize((size_t)n);
    for (int i = 0; i < n; ++i) g->m[i].ctx = ctxs[i], g->m[i].rank = i;
    if (n > 1) {
        std::string err;
        if (!nccl_load(&g->nc, err)) {
            delete g;
            return vo_fail(ctxs[0], VO_ERR_UNSUPPORTED, err);
        }
        std::vector<int> devs((size_t)n);
        std::vector<ncclComm_t> comms((size_t)n);
        for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
        ncclResult_t r = g->nc->commInitAll(comms.data(), n, devs.data());  // one process, every GPU of the group
        if (r != ncclSuccess) {
            const std::string msg = std::string("ncclCommInitAll: ") + g->nc->getErrorString(r);
            delete g;
            return vo_fail(ctxs[0], VO_ERR_NCCL, msg);
        }
        for (int i = 0; i < n; ++i) g->m[i].comm = comms[i];
    }
    int32_t rc = member_alloc(g);
    if (rc != VO_OK) {
        vo_group_destroy(g);
        return rc;
    }
    *out = g;
    return VO_OK;
}

int32_t vo_group_destroy(vo_group g) {
    if (!g) return VO_OK;
    for (Member& mb : g->m) {
        DeviceGuard dg(mb.ctx->device);
        cudaStreamSynchronize(mb.ctx->stream);
        if (mb.comm) g->nc->commDestroy(mb.comm);
        vo_dfree(mb.gbuf), vo_dfree(mb.scratch), cudaFreeHost(mb.pinned);
    }
    delete g;
    return VO_OK;
}

int32_t vo_group_world(vo_group g) { return g ? g->world : VO_ERR_BAD_ARG; }
int32_t vo_group_local_members(vo_group g) { return g ? (int32_t)g->m.size() : VO_ERR_BAD_ARG; }
int32_t vo_group_member_rank(vo_group g, int32_t member) { return (g && member >= 0 && member < (int32_t)g->m.size()) ? g->m[member].rank : VO_ERR_BAD_ARG; }
const char* vo_group_last_error(vo_group g) { return g ? g->err.c_str() : g_vo_tls_err.c_str(); }

int32_t vo_group_shard_range(int64_t n_total, int32_t rank, int32_t world, int64_t* lo, int64_t* hi) {
    if (n_total < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return VO_ERR_BAD_ARG;
    shard_range(n_total, rank, world, lo, hi);
    return VO_OK;
}

int32_t vo_group_nccl_version(void) {
    Nccl* nc = nullptr;
    std::string err;
    if (!nccl_load(&nc, err)) return vo_fail(nullptr, VO_ERR_UNSUPPORTED, err);
    int v = 0;
    nc->getVersion(&v);
    return v;
}

// The sharded ensemble -> one device ensemble [d][n_total] on the root. local[i] is member i's shard (SoA [d][n_i], n_i the
// length of shard_range(n_total, rank_i, world)). Component rows are sent one by one so that they land where the global
// SoA layout wants them: d x (world - 1) point-to-point transfers in one NCCL group, all in flight at once over NVSwitch.
int32_t vo_group_gather_device(vo_group g, const vo_ens* local, int64_t n_total, int32_t root, vo_ens* out) {
    if (!g || !local || n_total < 1 || root < 0 || root >= g->world) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_gather: bad argument");
    const int64_t d = local[0] ? local[0]->d : 0;
    for (size_t i = 0; i < g->m.size(); ++i) {
        int64_t lo, hi;
        shard_range(n_total, g->m[i].rank, g->world, &lo, &hi);
        if (!local[i] || local[i]->d != d || local[i]->n != hi - lo || local[i]->ctx != g->m[i].ctx)
            return g_fail(g, VO_ERR_SHAPE, "vo_group_gather: member " + std::to_string(i) + " must hand in its shard of ceil(N/G) trajectories on the group's ctx");
    }
    Member* rm = find_member(g, root);
    if (rm) {
        DeviceGuard dg(rm->ctx->device);
        if (rm->gbuf_elems < (size_t)(d * n_total)) {
            cudaStreamSynchronize(rm->ctx->stream);
            vo_dfree(rm->gbuf), rm->gbuf = nullptr, rm->gbuf_elems = 0;
            if (vo_dmalloc(&rm->gbuf, sizeof(double) * d * n_total) != cudaSuccess) return g_fail(g, VO_ERR_ALLOC, "vo_group_gather: gather buffer");
            rm->gbuf_elems = (size_t)(d * n_total);
        }
    }
    GroupScope gs;
    if (g->world > 1) VO_NCCL(g, gs.start(g->nc));
    for (size_t i = 0; i < g->m.size(); ++i) {
        Member& mb = g->m[i];
        DeviceGuard dg(mb.ctx->device);
        vo_touch(mb.ctx);
        int64_t lo, hi;
        shard_range(n_total, mb.rank, g->world, &lo, &hi);
        if (mb.rank == root) {
            if (hi > lo) {
                cudaError_t e = cudaMemcpy2DAsync(mb.gbuf + lo, sizeof(double) * n_total, local[i]->p, sizeof(double) * (hi - lo), sizeof(double) * (hi - lo), (size_t)d,
                                                  cudaMemcpyDeviceToDevice, mb.ctx->stream);
                if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_gather: ") + cudaGetErrorString(e));
            }
            for (int r = 0; r < g->world; ++r) {
                if (r == root) continue;
                int64_t rlo, rhi;
                shard_range(n_total, r, g->world, &rlo, &rhi);
                for (int64_t c = 0; c < d && rhi > rlo; ++c)
                    VO_NCCL(g, g->nc->recv(mb.gbuf + c * n_total + rlo, (size_t)(rhi - rlo), ncclDouble, r, mb.comm, mb.ctx->stream));
            }
        } else {
            for (int64_t c = 0; c < d && hi > lo; ++c)
                VO_NCCL(g, g->nc->send(local[i]->p + c * (hi - lo), (size_t)(hi - lo), ncclDouble, root, mb.comm, mb.ctx->stream));
        }
    }
    if (g->world > 1) VO_NCCL(g, gs.end());
    if (out) *out = nullptr;
    if (rm && out) return vo_ens_wrap(rm->ctx, rm->gbuf, d, n_total, out);
    return VO_OK;
}

// The general form of the final gather: rank r holds rows[r] trajectories (any sizes) and the root places them at row
// row_off[r] of a host array of host_n trajectories ([host_n][d] AoS or [d][host_n] SoA). Each shard travels as ONE message
// (its SoA block [d][rows_r]) into the root's gather buffer; the root then moves every block to the host with its own copy
// — AoS through a transposing kernel into a second staging area — so that a caller can gather the CHUNKS of a pipelined
// solve one by one, each straight into its place, while later chunks still integrate. Asynchronous: everything is enqueued
// on the members' streams; vo_group_sync waits for it.
__global__ void soa_block_to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos, int64_t d, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int64_t c = 0; c < d; ++c) aos[i * d + c] = soa[c * n + i];
}

int32_t vo_group_gather_placed(vo_group g, const vo_ens* local, int32_t root, const int64_t* rows, const int64_t* row_off, double* host_out, int32_t layout,
                               int64_t host_n) {
    if (!g || !local || !rows || !row_off || root < 0 || root >= g->world || (layout != VO_LAYOUT_AOS && layout != VO_LAYOUT_SOA))
        return g_fail(g, VO_ERR_BAD_ARG, "vo_group_gather_placed: bad argument");
    const int64_t d = local[0] ? local[0]->d : 0;
    int64_t total = 0;
    std::vector<int64_t> start((size_t)g->world);
    for (int r = 0; r < g->world; ++r) {
        if (rows[r] < 0 || row_off[r] < 0 || row_off[r] + rows[r] > host_n) return g_fail(g, VO_ERR_SHAPE, "vo_group_gather_placed: a shard does not fit the host array");
        start[r] = total, total += rows[r];
    }
    for (size_t i = 0; i < g->m.size(); ++i)
        if (!local[i] || local[i]->d != d || (rows[g->m[i].rank] > 0 && local[i]->n != rows[g->m[i].rank]) || local[i]->ctx->device != g->m[i].ctx->device)
            return g_fail(g, VO_ERR_SHAPE, "vo_group_gather_placed: member " + std::to_string(i) + " must hand in an ensemble of rows[rank] trajectories on its device");
    Member* rm = find_member(g, root);
    const bool aos = layout == VO_LAYOUT_AOS && d > 1;
    if (rm) {
        if (!host_out) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_gather_placed: the root needs a host buffer");
        DeviceGuard dg(rm->ctx->device);
        const size_t need = (size_t)(d * total) * (aos ? 2 : 1);  // second half: the AoS images
        if (rm->gbuf_elems < need) {
            cudaStreamSynchronize(rm->ctx->stream);
            vo_dfree(rm->gbuf), rm->gbuf = nullptr, rm->gbuf_elems = 0;
            if (vo_dmalloc(&rm->gbuf, sizeof(double) * need) != cudaSuccess) return g_fail(g, VO_ERR_ALLOC, "vo_group_gather_placed: gather buffer");
            rm->gbuf_elems = need;
        }
    }
    GroupScope gs;
    if (g->world > 1) VO_NCCL(g, gs.start(g->nc));
    for (size_t i = 0; i < g->m.size(); ++i) {
        Member& mb = g->m[i];
        DeviceGuard dg(mb.ctx->device);
        vo_touch(mb.ctx);
        const int64_t n_i = rows[mb.rank];
        if (mb.rank == root) {
            if (n_i > 0) {
                cudaError_t e = cudaMemcpyAsync(mb.gbuf + d * start[root], local[i]->p, sizeof(double) * d * n_i, cudaMemcpyDeviceToDevice, mb.ctx->stream);
                if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_gather_placed: ") + cudaGetErrorString(e));
            }
            for (int r = 0; r < g->world; ++r)
                if (r != root && rows[r] > 0) VO_NCCL(g, g->nc->recv(mb.gbuf + d * start[r], (size_t)(d * rows[r]), ncclDouble, r, mb.comm, mb.ctx->stream));
        } else if (n_i > 0) {
            VO_NCCL(g, g->nc->send(local[i]->p, (size_t)(d * n_i), ncclDouble, root, mb.comm, mb.ctx->stream));
        }
    }
    if (g->world > 1) VO_NCCL(g, gs.end());
    if (rm) {
        DeviceGuard dg(rm->ctx->device);
        cudaStream_t st = rm->ctx->stream;
        for (int r = 0; r < g->world; ++r) {
            const int64_t n_r = rows[r];
            if (n_r == 0) continue;
            const double* blk = rm->gbuf + d * start[r];
            cudaError_t e;
            if (aos) {
                double* img = rm->gbuf + d * total + d * start[r];
                soa_block_to_aos_kernel<<<(unsigned)ceil_div(n_r, 256), 256, 0, st>>>(blk, img, d, n_r);
                rm->ctx->launches++;
                e = cudaMemcpyAsync(host_out + row_off[r] * d, img, sizeof(double) * d * n_r, cudaMemcpyDeviceToHost, st);
            } else {
                e = cudaMemcpy2DAsync(host_out + row_off[r], sizeof(double) * host_n, blk, sizeof(double) * n_r, sizeof(double) * n_r, (size_t)d, cudaMemcpyDeviceToHost, st);
            }
            if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_gather_placed: ") + cudaGetErrorString(e));
        }
    }
    return VO_OK;
}

// Round-robin sharding: rank r holds the trajectories j = r, r + G, r + 2G, ... of a block of `tot` consecutive trajectories
// (rows[r] = ceil((tot - r) / G) of them), the static interleave that balances an adaptive ensemble whose cost varies smoothly
// along the trajectory index. The root receives every shard as one message, interleaves them on the device
// into the block's natural order and moves the block to rows [row0, row0 + tot) of the host array with ONE copy. Asynchronous
// like vo_group_gather_placed.
struct InterleaveArgs {
    int64_t start[16], rows[16];
    int G;
};
__global__ void interleave_kernel(const double* __restrict__ gbuf, double* __restrict__ img, InterleaveArgs a, int64_t d, int64_t tot, int aos) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= tot) return;
    const int r = (int)(j % a.G);
    const int64_t i = j / a.G, n_r = a.rows[r];
    const double* blk = gbuf + d * a.start[r];
    for (int64_t c = 0; c < d; ++c) img[aos ? j * d + c : c * tot + j] = blk[c * n_r + i];
}

int32_t vo_group_gather_interleaved(vo_group g, const vo_ens* local, int32_t root, int64_t tot, int64_t row0, double* host_out, int32_t layout, int64_t host_n) {
    if (!g || !local || root < 0 || root >= g->world || g->world > 16 || tot < 0 || row0 < 0 || row0 + tot > host_n || (layout != VO_LAYOUT_AOS && layout != VO_LAYOUT_SOA))
        return g_fail(g, VO_ERR_BAD_ARG, "vo_group_gather_interleaved: bad argument");
    const int G = g->world;
    const int64_t d = local[0] ? local[0]->d : 0;
    InterleaveArgs ia;
    std::memset(&ia, 0, sizeof ia);
    ia.G = G;
    int64_t total = 0;
    for (int r = 0; r < G; ++r) ia.rows[r] = tot > r ? (tot - r + G - 1) / G : 0, ia.start[r] = total, total += ia.rows[r];
    for (size_t i = 0; i < g->m.size(); ++i)
        if (!local[i] || local[i]->d != d || (ia.rows[g->m[i].rank] > 0 && local[i]->n != ia.rows[g->m[i].rank]) || local[i]->ctx->device != g->m[i].ctx->device)
            return g_fail(g, VO_ERR_SHAPE, "vo_group_gather_interleaved: member " + std::to_string(i) + " must hand in ceil((tot - rank) / G) trajectories on its device");
    Member* rm = find_member(g, root);
    if (rm) {
        if (!host_out) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_gather_interleaved: the root needs a host buffer");
        DeviceGuard dg(rm->ctx->device);
        const size_t need = (size_t)(d * total) * 2;
        if (rm->gbuf_elems < need) {
            cudaStreamSynchronize(rm->ctx->stream);
            vo_dfree(rm->gbuf), rm->gbuf = nullptr, rm->gbuf_elems = 0;
            if (vo_dmalloc(&rm->gbuf, sizeof(double) * need) != cudaSuccess) return g_fail(g, VO_ERR_ALLOC, "vo_group_gather_interleaved: gather buffer");
            rm->gbuf_elems = need;
        }
    }
    GroupScope gs;
    if (G > 1) VO_NCCL(g, gs.start(g->nc));
    for (size_t i = 0; i < g->m.size(); ++i) {
        Member& mb = g->m[i];
        DeviceGuard dg(mb.ctx->device);
        vo_touch(mb.ctx);
        const int64_t n_i = ia.rows[mb.rank];
        if (mb.rank == root) {
            if (n_i > 0) {
                cudaError_t e = cudaMemcpyAsync(mb.gbuf + d * ia.start[root], local[i]->p, sizeof(double) * d * n_i, cudaMemcpyDeviceToDevice, mb.ctx->stream);
                if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_gather_interleaved: ") + cudaGetErrorString(e));
            }
            for (int r = 0; r < G; ++r)
                if (r != root && ia.rows[r] > 0) VO_NCCL(g, g->nc->recv(mb.gbuf + d * ia.start[r], (size_t)(d * ia.rows[r]), ncclDouble, r, mb.comm, mb.ctx->stream));
        } else if (n_i > 0) {
            VO_NCCL(g, g->nc->send(local[i]->p, (size_t)(d * n_i), ncclDouble, root, mb.comm, mb.ctx->stream));
        }
    }
    if (G > 1) VO_NCCL(g, gs.end());
    if (rm && tot > 0) {
        DeviceGuard dg(rm->ctx->device);
        cudaStream_t st = rm->ctx->stream;
        double* img = rm->gbuf + d * total;
        const int aos = (layout == VO_LAYOUT_AOS) ? 1 : 0;
        interleave_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(rm->gbuf, img, ia, d, tot, aos);
        rm->ctx->launches++;
        cudaError_t e = aos ? cudaMemcpyAsync(host_out + row0 * d, img, sizeof(double) * d * tot, cudaMemcpyDeviceToHost, st)
                            : cudaMemcpy2DAsync(host_out + row0, sizeof(double) * host_n, img, sizeof(double) * tot, sizeof(double) * tot, (size_t)d, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_gather_interleaved: ") + cudaGetErrorString(e));
    }
    return VO_OK;
}

int32_t vo_group_sync(vo_group g) {
    if (!g) return VO_ERR_BAD_ARG;
    for (Member& mb : g->m) {
        DeviceGuard dg(mb.ctx->device);
        cudaError_t e = cudaStreamSynchronize(mb.ctx->stream);
        if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_sync: ") + cudaGetErrorString(e));
    }
    return VO_OK;
}

int32_t vo_group_gather(vo_group g, const vo_ens* local, int64_t n_total, int32_t root, double* host_out, int32_t layout) {
    vo_ens whole = nullptr;
    int32_t rc = vo_group_gather_device(g, local, n_total, root, &whole);
    if (rc != VO_OK) return rc;
    if (whole) {  // this process holds the root
        if (!host_out) {
            vo_ens_destroy(whole);
            return g_fail(g, VO_ERR_BAD_ARG, "vo_group_gather: the root needs a host buffer");
        }
        rc = vo_ens_download(whole, host_out, layout);  // one device-to-host copy of the whole ensemble (synchronous on return)
        vo_ens_destroy(whole);
    }
    return rc;
}

// Inverse of the gather for the initial condition: the root uploads the whole ensemble once and the shards travel over NVLink.
int32_t vo_group_scatter(vo_group g, const double* host_in, int32_t layout, int64_t d, int64_t n_total, int32_t root, const vo_ens* local_out) {
    if (!g || !local_out || n_total < 1 || d < 1 || root < 0 || root >= g->world) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_scatter: bad argument");
    for (size_t i = 0; i < g->m.size(); ++i) {
        int64_t lo, hi;
        shard_range(n_total, g->m[i].rank, g->world, &lo, &hi);
        if (!local_out[i] || local_out[i]->d != d || local_out[i]->n != hi - lo || local_out[i]->ctx != g->m[i].ctx)
            return g_fail(g, VO_ERR_SHAPE, "vo_group_scatter: member " + std::to_string(i) + " must hand in an ensemble of its shard's shape on the group's ctx");
    }
    Member* rm = find_member(g, root);
    if (rm) {
        if (!host_in) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_scatter: the root needs the host buffer");
        DeviceGuard dg(rm->ctx->device);
        if (rm->gbuf_elems < (size_t)(d * n_total)) {
            cudaStreamSynchronize(rm->ctx->stream);
            vo_dfree(rm->gbuf), rm->gbuf = nullptr, rm->gbuf_elems = 0;
            if (vo_dmalloc(&rm->gbuf, sizeof(double) * d * n_total) != cudaSuccess) return g_fail(g, VO_ERR_ALLOC, "vo_group_scatter: staging buffer");
            rm->gbuf_elems = (size_t)(d * n_total);
        }
        vo_ens whole = nullptr;
        int32_t rc = vo_ens_wrap(rm->ctx, rm->gbuf, d, n_total, &whole);
        if (rc == VO_OK) rc = vo_ens_upload(whole, host_in, layout);
        vo_ens_destroy(whole);
        if (rc != VO_OK) return rc;
    }
    GroupScope gs;
    if (g->world > 1) VO_NCCL(g, gs.start(g->nc));
    for (size_t i = 0; i < g->m.size(); ++i) {
        Member& mb = g->m[i];
        DeviceGuard dg(mb.ctx->device);
        vo_touch(mb.ctx);
        int64_t lo, hi;
        shard_range(n_total, mb.rank, g->world, &lo, &hi);
        if (mb.rank == root) {
            if (hi > lo) {
                cudaError_t e = cudaMemcpy2DAsync(local_out[i]->p, sizeof(double) * (hi - lo), mb.gbuf + lo, sizeof(double) * n_total, sizeof(double) * (hi - lo), (size_t)d,
                                                  cudaMemcpyDeviceToDevice, mb.ctx->stream);
                if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_scatter: ") + cudaGetErrorString(e));
            }
            for (int r = 0; r < g->world; ++r) {
                if (r == root) continue;
                int64_t rlo, rhi;
                shard_range(n_total, r, g->world, &rlo, &rhi);
                for (int64_t c = 0; c < d && rhi > rlo; ++c)
                    VO_NCCL(g, g->nc->send(mb.gbuf + c * n_total + rlo, (size_t)(rhi - rlo), ncclDouble, r, mb.comm, mb.ctx->stream));
            }
        } else {
            for (int64_t c = 0; c < d && hi > lo; ++c)
                VO_NCCL(g, g->nc->recv(local_out[i]->p + c * (hi - lo), (size_t)(hi - lo), ncclDouble, root, mb.comm, mb.ctx->stream));
        }
    }
    if (g->world > 1) VO_NCCL(g, gs.end());
    for (Member& mb : g->m) {
        DeviceGuard dg(mb.ctx->device);
        cudaError_t e = cudaStreamSynchronize(mb.ctx->stream);
        if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_scatter: ") + cudaGetErrorString(e));
    }
    return VO_OK;
}

// In-place all-reduce of n doubles per member (host values in, reduced values out). op: 0 sum, 1 max, 2 min.
int32_t vo_group_allreduce(vo_group g, double* host_inout /* [members][n] */, int32_t n, int32_t op) {
    if (!g || !host_inout || n < 1 || n > 32 || op < 0 || op > 2) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_allreduce: bad argument (n <= 32)");
    const size_t nm = g->m.size();
    if (g->world == 1) return VO_OK;
    for (size_t i = 0; i < nm; ++i) {
        Member& mb = g->m[i];
        DeviceGuard dg(mb.ctx->device);
        std::memcpy(mb.pinned, host_inout + i * n, sizeof(double) * n);
        cudaMemcpyAsync(mb.scratch, mb.pinned, sizeof(double) * n, cudaMemcpyHostToDevice, mb.ctx->stream);
    }
    GroupScope gs;
    VO_NCCL(g, gs.start(g->nc));
    for (Member& mb : g->m) {
        DeviceGuard dg(mb.ctx->device);
        VO_NCCL(g, g->nc->allReduce(mb.scratch, mb.scratch, (size_t)n, ncclDouble, op == 0 ? ncclSum : op == 1 ? ncclMax : ncclMin, mb.comm, mb.ctx->stream));
    }
    VO_NCCL(g, gs.end());
    for (size_t i = 0; i < nm; ++i) {
        Member& mb = g->m[i];
        DeviceGuard dg(mb.ctx->device);
        cudaMemcpyAsync(mb.pinned, mb.scratch, sizeof(double) * n, cudaMemcpyDeviceToHost, mb.ctx->stream);
        cudaError_t e = cudaStreamSynchronize(mb.ctx->stream);
        if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_allreduce: ") + cudaGetErrorString(e));
        std::memcpy(host_inout + i * n, mb.pinned, sizeof(double) * n);
    }
    return VO_OK;
}

// Counters of the whole sharded ensemble: every member reduces its own solver on its device (one small kernel), then two
// all-reduces (integer sums; max of (t_max, -t_min)). Every rank receives the totals.
int32_t vo_group_reduce_stats(vo_group g, const vo_solver* local, vo_group_stats* out) {
    if (!g || !local || !out) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_reduce_stats: bad argument");
    for (size_t i = 0; i < g->m.size(); ++i) {
        Member& mb = g->m[i];
        if (!local[i]) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_reduce_stats: NULL solver");
        DeviceGuard dg(mb.ctx->device);
        int32_t rc = vo_solver_local_stats(local[i], (unsigned long long*)mb.scratch, (double*)((char*)mb.scratch + 64));
        if (rc != VO_OK) return rc;
    }
    if (g->world > 1) {
        GroupScope gs;
        VO_NCCL(g, gs.start(g->nc));
        for (Member& mb : g->m) {
            DeviceGuard dg(mb.ctx->device);
            VO_NCCL(g, g->nc->allReduce(mb.scratch, mb.scratch, 6, ncclUint64, ncclSum, mb.comm, mb.ctx->stream));
            VO_NCCL(g, g->nc->allReduce((char*)mb.scratch + 64, (char*)mb.scratch + 64, 2, ncclDouble, ncclMax, mb.comm, mb.ctx->stream));
        }
        VO_NCCL(g, gs.end());
    }
    Member& m0 = g->m[0];
    DeviceGuard dg(m0.ctx->device);
    cudaMemcpyAsync(m0.pinned, m0.scratch, 128, cudaMemcpyDeviceToHost, m0.ctx->stream);
    for (Member& mb : g->m) {
        DeviceGuard dg2(mb.ctx->device);
        cudaError_t e = cudaStreamSynchronize(mb.ctx->stream);
        if (e != cudaSuccess) return g_fail(g, VO_ERR_CUDA, std::string("vo_group_reduce_stats: ") + cudaGetErrorString(e));
    }
    const unsigned long long* s = (const unsigned long long*)m0.pinned;
    const double* mm = (const double*)((const char*)m0.pinned + 64);
    out->accepted = (int64_t)s[0], out->rejected = (int64_t)s[1], out->n_traj = (int64_t)s[2], out->n_done = (int64_t)s[3];
    out->n_nonfinite = (int64_t)s[4], out->n_stuck = (int64_t)s[5];
    out->t_max = mm[0], out->t_min = -mm[1];
    return VO_OK;
}

// `while let Ok(_) = solver.step() {}` on every shard, then the reduction of the counters. The shards of one process are
// driven round-robin from this one host thread; nothing is exchanged while stepping.
int32_t vo_group_run(vo_group g, const vo_solver* local, int32_t adaptive, int64_t max_calls, vo_group_stats* out) {
    if (!g || !local) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_run: bad argument");
    for (size_t i = 0; i < g->m.size(); ++i) {
        if (!local[i]) return g_fail(g, VO_ERR_BAD_ARG, "vo_group_run: NULL solver");
        vo_step_result res;
        int32_t rc = vo_run(local[i], adaptive, max_calls, &res);
        if (rc != VO_OK) return rc;
    }
    return out ? vo_group_reduce_stats(g, local, out) : VO_OK;
}

}  // extern "C"
