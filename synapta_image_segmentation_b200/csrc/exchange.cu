// exchange.cu -- the path's single exchange step: cross-page duplicate-figure removal over all ranks.
//
// Every rank hashes the candidate regions of its own pages (phash.cu); then ONE ncclAllGather over NVLink moves each
// rank's fixed-capacity block { count, (key, hash) x capacity } to everybody (about 150 KB at 8,000 pages: latency-bound,
// so the collective is NCCL's and not a hand-written peer-memory kernel), and every rank runs the same kernels on the
// gathered blocks: compact the valid pairs, rank every pair by its key and test it against all pairs with a smaller key
// (Hamming distance <= t) in one all-pairs pass, scatter by rank.  The result -- keys in ascending order with their keep
// flags -- is identical on all ranks without a second collective.  No torch op, no host synchronisation inside.
// The reference's nearest analogue is the id-dedup on md5 (pdf_image_segmentation.py:3886-3887).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process -- the copy torch.distributed loaded --
// else the system one), so libsynseg.so has no link-time dependency on it and a single-GPU process never loads it.
#include <dlfcn.h>
#include <nccl.h>

#include "internal.cuh"

namespace {

struct NcclApi {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
};
NcclApi g_nccl = {};

int nccl_load()
{
    if (g_nccl.lib) return SYNSEG_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);         // the copy already mapped (torch's), if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { synseg_set_error("NCCL: cannot load libnccl.so.2 (%s)", dlerror()); return SYNSEG_E_INVALID; }
    NcclApi a = {};
    a.lib = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString) {
        synseg_set_error("NCCL: libnccl.so.2 lacks a required symbol"); return SYNSEG_E_INVALID;
    }
    g_nccl = a;
    return SYNSEG_OK;
}

int nccl_check(ncclResult_t r, const char *what)
{
    if (r == ncclSuccess) return SYNSEG_OK;
    synseg_set_error("NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    return SYNSEG_E_CUDA;
}

typedef unsigned long long u64;

// block = { count, 0, (key, hash) x capacity }
__global__ void __launch_bounds__(256) exchange_pack_kernel(const u64 *hashes, const u64 *keys, const int32_t *count, int capacity, u64 *block)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n = min(max(*count, 0), capacity);
    if (i == 0) { block[0] = (u64)n; block[1] = 0; }
    if (i < n) { block[2 + 2 * (int64_t)i] = keys[i]; block[3 + 2 * (int64_t)i] = hashes[i]; }
}

// gathered blocks -> dense (key, hash) arrays; n_total = sum of the counts; rank / keep scratch preset
__global__ void __launch_bounds__(256) exchange_compact_kernel(const u64 *blocks, int world, int capacity, u64 *dkeys, u64 *dhash, int32_t *rank,
                                                               uint8_t *dkeep, int32_t *n_total)
{
    const int r = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    const int64_t stride = 2 + 2 * (int64_t)capacity;
    int off = 0, total = 0;
    for (int q = 0; q < world; ++q) { const int c = (int)blocks[q * stride]; if (q < r) off += c; total += c; }
    if (r == 0 && j == 0) *n_total = total;
    const int cnt = (int)blocks[r * stride];
    if (j >= cnt) return;
    const u64 *b = blocks + r * stride + 2 + 2 * (int64_t)j;
    SS_DEVICE_ASSERT(cnt <= capacity && off + j < world * capacity);
    dkeys[off + j] = b[0]; dhash[off + j] = b[1];
    rank[off + j] = 0; dkeep[off + j] = 1;
}

// All pairs, j cut into slices (blockIdx.y) staged in shared memory: rank[i] += #{j in slice: key[j] < key[i]};
// keep[i] = 0 when one of those j is within max_hamming of hash[i].
constexpr int EX_TILE = 1024;
__global__ void __launch_bounds__(256) exchange_pairs_kernel(const u64 *dkeys, const u64 *dhash, const int32_t *n_total, int max_hamming,
                                                             int32_t *rank, uint8_t *dkeep)
{
    __shared__ u64 sk[EX_TILE], sh[EX_TILE];
    const int n = *n_total;
    const int j0 = blockIdx.y * EX_TILE;
    if (j0 >= n || (int)blockIdx.x * 256 >= n) return;                // uniform per CTA
    const int j1 = min(j0 + EX_TILE, n);
    for (int j = j0 + threadIdx.x; j < j1; j += 256) { sk[j - j0] = dkeys[j]; sh[j - j0] = dhash[j]; }
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const u64 k = dkeys[i], h = dhash[i];
    int smaller = 0;
    bool dup = false;
    for (int j = 0; j < j1 - j0; ++j) {
        const bool lt = sk[j] < k;
        smaller += lt;
        dup |= lt && (__popcll(sh[j] ^ h) <= max_hamming);
    }
    if (smaller) atomicAdd(&rank[i], smaller);
    if (dup) dkeep[i] = 0;
}

__global__ void __launch_bounds__(256) exchange_scatter_kernel(const u64 *dkeys, const u64 *dhash, const int32_t *rank, const uint8_t *dkeep,
                                                               const int32_t *n_total, u64 *all_keys, u64 *all_hashes, uint8_t *keep)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= *n_total) return;
    const int r = rank[i];
    SS_DEVICE_ASSERT(r >= 0 && r < *n_total);
    all_keys[r] = dkeys[i];
    if (all_hashes) all_hashes[r] = dhash[i];
    keep[r] = dkeep[i];
}

}  // namespace

void comm_release(synseg_ctx *ctx)
{
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr; ctx->comm_world = 1; ctx->comm_rank = 0;
}

extern "C" SYNSEG_EXPORT int synseg_comm_unique_id(uint8_t *id128)
{
    if (!id128) { synseg_set_error("synseg_comm_unique_id: NULL argument"); return SYNSEG_E_INVALID; }
    SS_TRY(nccl_load());
    ncclUniqueId id;
    SS_TRY(nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId"));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, sizeof(id));
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_comm_init(synseg_ctx *ctx, const uint8_t *id128, int32_t rank, int32_t world)
{
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) { synseg_set_error("synseg_comm_init: bad arguments"); return SYNSEG_E_INVALID; }
    DeviceScope scope(ctx->device);
    comm_release(ctx);
    if (world == 1) return SYNSEG_OK;
    SS_TRY(nccl_load());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    SS_TRY(nccl_check(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank"));
    ctx->comm = comm; ctx->comm_world = world; ctx->comm_rank = rank;
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_comm_destroy(synseg_ctx *ctx)
{
    if (!ctx) return SYNSEG_OK;
    DeviceScope scope(ctx->device);
    cudaDeviceSynchronize();
    comm_release(ctx);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_comm_info(const synseg_ctx *ctx, int32_t *rank, int32_t *world, int32_t *nccl_version)
{
    if (!ctx) { synseg_set_error("synseg_comm_info: ctx is NULL"); return SYNSEG_E_INVALID; }
    if (rank) *rank = ctx->comm_rank;
    if (world) *world = ctx->comm ? ctx->comm_world : 1;
    if (nccl_version) { int v = 0; if (g_nccl.GetVersion) g_nccl.GetVersion(&v); *nccl_version = v; }
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_dedup_exchange(synseg_ctx *ctx, const uint64_t *hashes, const uint64_t *keys, const int32_t *count,
                                                   int32_t capacity, int32_t max_hamming, uint64_t *all_keys, uint64_t *all_hashes, uint8_t *keep,
                                                   int32_t *n_total, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_dedup_exchange: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (!hashes || !keys || !count || !all_keys || !keep || !n_total || capacity < 1 || capacity > (1 << 22)) {
        synseg_set_error("synseg_dedup_exchange: bad arguments"); return SYNSEG_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int world = ctx->comm ? ctx->comm_world : 1;
    const size_t block = 2 + 2 * (size_t)capacity;                    // u64 words per rank
    const size_t cap_total = (size_t)world * capacity;
    const size_t need = sizeof(u64) * (block * (world + 1) + 2 * cap_total) + sizeof(int32_t) * cap_total + cap_total + 8 * 256;
    SS_TRY(arena_ensure(ctx, need));
    arena_begin(ctx);
    void *p;
    SS_TRY(arena_alloc(ctx, sizeof(u64) * block, &p, st)); u64 *send = (u64 *)p;
    SS_TRY(arena_alloc(ctx, sizeof(u64) * block * world, &p, st)); u64 *recv = (u64 *)p;
    SS_TRY(arena_alloc(ctx, sizeof(u64) * cap_total, &p, st)); u64 *dkeys = (u64 *)p;
    SS_TRY(arena_alloc(ctx, sizeof(u64) * cap_total, &p, st)); u64 *dhash = (u64 *)p;
    SS_TRY(arena_alloc(ctx, sizeof(int32_t) * cap_total, &p, st)); int32_t *rank = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, cap_total, &p, st)); uint8_t *dkeep = (uint8_t *)p;
    exchange_pack_kernel<<<cdiv(capacity, 256), 256, 0, st>>>((const u64 *)hashes, (const u64 *)keys, count, capacity, world > 1 ? send : recv);
    SS_LAUNCH_CHECK(ctx, "exchange_pack", st);
    if (world > 1) SS_TRY(nccl_check(g_nccl.AllGather(send, recv, block, ncclUint64, (ncclComm_t)ctx->comm, st), "ncclAllGather"));
    exchange_compact_kernel<<<dim3(cdiv(capacity, 256), world), 256, 0, st>>>(recv, world, capacity, dkeys, dhash, rank, dkeep, n_total);
    SS_LAUNCH_CHECK(ctx, "exchange_compact", st);
    exchange_pairs_kernel<<<dim3(cdiv(cap_total, 256), cdiv(cap_total, EX_TILE)), 256, 0, st>>>(dkeys, dhash, n_total, max_hamming, rank, dkeep);
    SS_LAUNCH_CHECK(ctx, "exchange_pairs", st);
    exchange_scatter_kernel<<<cdiv(cap_total, 256), 256, 0, st>>>(dkeys, dhash, rank, dkeep, n_total, (u64 *)all_keys, (u64 *)all_hashes, keep);
    SS_LAUNCH_CHECK(ctx, "exchange_scatter", st);
    return SYNSEG_OK;
}
