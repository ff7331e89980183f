// ccl.cu -- 8-connected component labelling with fused statistics, bit-exact with
// cv2.connectedComponentsWithStats(mask, 8, CV_32S) including cv2's label numbering.
//
// No reference call site for the labelling itself (north-star primitive); it also
// replaces findContours(RETR_EXTERNAL)+boundingRect at pdf_image_segmentation.py:1403-1404.
// The same union-find core implements Canny's hysteresis (canny.cu).
//
// Algorithm: block-based union-find.  One thread owns a 2x2 pixel block (all foreground pixels of a
// block are mutually 8-connected, so one label per block suffices).  init: L[b] = b (foreground) or
// -1; merge: lock-free union (atomicMin towards the smaller block index) with the four predecessor
// blocks whose pixels touch; compress: L[b] = root.  The root of a component is therefore its
// smallest block index = the first block in 2x2-block raster order, which is exactly the order in
// which cv2's block-based algorithm numbers components; final labels are 1 + (rank of the root
// among roots), obtained with a chunked scan over the block array.
// Statistics (bbox, area, coordinate sums) are reduced hierarchically: redux.sync over lanes that
// share a label -> a small per-CTA shared-memory cache keyed by label -> one global atomic set per
// CTA and label.  Centroids are sum/area in f64 like cv2.
//
// Roofline: HBM-bound; algorithmic bytes 5 per pixel with a label image (1 read + 4 written),
// 1 per pixel without (mask read; the block-label scratch is 1 B/px and L2-resident per page).
#include "internal.cuh"

namespace {

struct MaskAcc {
    const uint8_t *p; int64_t rs; int64_t bs;      // u8 plane (p != nullptr)
    const uint32_t *bits; int wpr; int64_t wbs;    // bit plane otherwise
    int width, height;
};

__device__ __forceinline__ bool fg_at(const MaskAcc &m, int img, int x, int y)
{
    if (m.p) return __ldg(m.p + img * m.bs + y * m.rs + x) != 0;
    return (__ldg(m.bits + img * m.wbs + (int64_t)y * m.wpr + (x >> 5)) >> (x & 31)) & 1u;
}

// 2x2 block pixel presence: bit0 (r,c), bit1 (r,c+1), bit2 (r+1,c), bit3 (r+1,c+1); in-image only.
__device__ __forceinline__ uint32_t block_px(const MaskAcc &m, int img, int c, int r)
{
    uint32_t v = 0;
    const bool x1 = c + 1 < m.width, y1 = r + 1 < m.height;
    if (m.p) {
        const uint8_t *row = m.p + img * m.bs + r * m.rs + c;
        v |= (__ldg(row) != 0) ? 1u : 0u;
        if (x1) v |= (__ldg(row + 1) != 0) ? 2u : 0u;
        if (y1) {
            v |= (__ldg(row + m.rs) != 0) ? 4u : 0u;
            if (x1) v |= (__ldg(row + m.rs + 1) != 0) ? 8u : 0u;
        }
    } else {
        // c is even, so both pixels of a row live in the same word
        const uint32_t *w = m.bits + img * m.wbs + (int64_t)r * m.wpr + (c >> 5);
        uint32_t a = (__ldg(w) >> (c & 31)) & 3u;
        if (!x1) a &= 1u;
        v = a;
        if (y1) {
            uint32_t b = (__ldg(w + m.wpr) >> (c & 31)) & 3u;
            if (!x1) b &= 1u;
            v |= b << 2;
        }
    }
    return v;
}

__device__ __forceinline__ int32_t uf_find(const int32_t *L, int32_t i)
{
    // L2-coherent loads: parents only ever decrease, so a stale value is still an ancestor
    int32_t p = __ldcg(L + i);
    while (p != i) { i = p; p = __ldcg(L + i); }
    return i;
}

__device__ __forceinline__ void uf_union(int32_t *L, int32_t a, int32_t b)
{
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { int32_t old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { int32_t old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

struct CclGeom { int bw, bh; int nblk; int64_t bper; };   // blocks per row / column / image; padded stride

__global__ void __launch_bounds__(256) ccl_init_kernel(MaskAcc m, CclGeom g, int32_t *L)
{
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y, img = blockIdx.z;
    if (bx >= g.bw || by >= g.bh) return;
    const int32_t b = by * g.bw + bx;
    L[img * g.bper + b] = block_px(m, img, 2 * bx, 2 * by) ? b : -1;
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(MaskAcc m, CclGeom g, int32_t *Lall)
{
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y, img = blockIdx.z;
    if (bx >= g.bw || by >= g.bh) return;
    const int c = 2 * bx, r = 2 * by;
    const uint32_t px = block_px(m, img, c, r);
    if (!px) return;
    // 4x4 window mask, bit = 4*wy + wx with window origin (r-1, c-1)
    uint32_t P = 0;
    if (px & 1u) P |= 0x777u;
    if (px & 2u) P |= 0x777u << 1;
    if (px & 4u) P |= 0x777u << 4;
    if (px & 8u) P |= 0x777u << 5;
    if (c == 0) P &= 0xEEEEu;
    if (c + 1 >= m.width) P &= 0x3333u;
    else if (c + 2 >= m.width) P &= 0x7777u;
    if (r == 0) P &= 0xFFF0u;
    int32_t *L = Lall + img * g.bper;
    const int32_t b = by * g.bw + bx;
    if ((P & 0x1u) && fg_at(m, img, c - 1, r - 1)) uf_union(L, b, b - g.bw - 1);
    if (((P & 0x2u) && fg_at(m, img, c, r - 1)) || ((P & 0x4u) && fg_at(m, img, c + 1, r - 1))) uf_union(L, b, b - g.bw);
    if ((P & 0x8u) && fg_at(m, img, c + 2, r - 1)) uf_union(L, b, b - g.bw + 1);
    if (((P & 0x10u) && fg_at(m, img, c - 1, r)) || ((P & 0x100u) && r + 1 < m.height && fg_at(m, img, c - 1, r + 1)))
        uf_union(L, b, b - 1);
}

__global__ void __launch_bounds__(256) ccl_compress_kernel(int64_t n, int32_t *L, int64_t bper, int nblk)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int64_t img = i / bper;
    int32_t *Li = L + img * bper;
    const int32_t b = (int32_t)(i - img * bper);
    if (b >= nblk) return;          // padding entries of the per-image stride are never initialised
    const int32_t v = Li[b];
    if (v >= 0 && v != b) Li[b] = uf_find(Li, v);
}

// ---- numbering: chunked scan over the block array -------------------------------------------
constexpr int CHUNK = 1024;   // block entries per CTA (256 threads x 4)

__device__ __forceinline__ int count_roots4(const int32_t *L, int nblk, int base, int t, bool flag[4])
{
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int idx = base + 4 * t + j;
        flag[j] = (idx < nblk) && (L[idx] == idx);
        cnt += flag[j];
    }
    return cnt;
}

__global__ void __launch_bounds__(256) ccl_count_kernel(const int32_t *Lall, int64_t bper, int nblk, int nchunks, int32_t *chunk_cnt)
{
    const int img = blockIdx.y, chunk = blockIdx.x;
    const int32_t *L = Lall + img * bper;
    bool f[4];
    int c = count_roots4(L, nblk, chunk * CHUNK, threadIdx.x, f);
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < 8; ++i) s += ws[i];
        chunk_cnt[img * nchunks + chunk] = s;
    }
}

// one CTA per image: exclusive scan of the chunk counts (in place), total -> n_roots[img]
__global__ void __launch_bounds__(256) ccl_scan_kernel(int nchunks, int32_t *chunk_cnt, int32_t *n_roots)
{
    const int img = blockIdx.x;
    int32_t *c = chunk_cnt + (int64_t)img * nchunks;
    __shared__ int ws[8];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nchunks; base += 256) {
        const int i = base + threadIdx.x;
        const int v = (i < nchunks) ? c[i] : 0;
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int nb = __shfl_up_sync(0xffffffffu, s, d); if ((threadIdx.x & 31) >= d) s += nb; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += ws[w];
        const int excl = carry + woff + s - v;
        if (i < nchunks) c[i] = excl;
        __syncthreads();
        if (threadIdx.x == 255) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_roots[img] = carry;
}

// roots get L[root] = -(label) - 1 with label = 1 + rank
__global__ void __launch_bounds__(256) ccl_assign_kernel(int32_t *Lall, int64_t bper, int nblk, int nchunks, const int32_t *chunk_off)
{
    const int img = blockIdx.y, chunk = blockIdx.x;
    int32_t *L = Lall + img * bper;
    bool f[4];
    const int cnt = count_roots4(L, nblk, chunk * CHUNK, threadIdx.x, f);
    int s = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int nb = __shfl_up_sync(0xffffffffu, s, d); if ((threadIdx.x & 31) >= d) s += nb; }
    __shared__ int ws[8];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += ws[w];
    int rank = chunk_off[img * nchunks + chunk] + woff + s - cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (f[j]) { L[chunk * CHUNK + 4 * threadIdx.x + j] = -(rank + 1) - 1; ++rank; }
}

// ---- statistics ---------------------------------------------------------------------------
struct StatAcc {          // SoA accumulators, [batch][cap]
    int32_t *minx, *miny, *maxx, *maxy, *area;
    unsigned long long *sumx, *sumy;
    int cap;
};

__global__ void stats_init_kernel(StatAcc a, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a.minx[i] = 0x7fffffff; a.miny[i] = 0x7fffffff; a.maxx[i] = -1; a.maxy[i] = -1; a.area[i] = 0;
    a.sumx[i] = 0; a.sumy[i] = 0;
}

constexpr int NSLOT = 16;
struct SlotCache {
    int key[NSLOT];
    int minx[NSLOT], miny[NSLOT], maxx[NSLOT], maxy[NSLOT], area[NSLOT];
    unsigned int sumx[NSLOT], sumy[NSLOT];
};

struct Contrib { int minx, miny, maxx, maxy, area; unsigned int sumx, sumy; };

// Lanes in `peers` share `key`: reduce their contributions; the leader adds into the CTA cache or global.
__device__ __forceinline__ void accumulate_group(unsigned peers, int key, Contrib v, SlotCache &sc, const StatAcc &a, int64_t img_off)
{
    v.minx = __reduce_min_sync(peers, v.minx); v.miny = __reduce_min_sync(peers, v.miny);
    v.maxx = __reduce_max_sync(peers, v.maxx); v.maxy = __reduce_max_sync(peers, v.maxy);
    v.area = __reduce_add_sync(peers, v.area);
    v.sumx = __reduce_add_sync(peers, v.sumx); v.sumy = __reduce_add_sync(peers, v.sumy);
    const int lane = threadIdx.x + threadIdx.y * 32;
    if ((lane & 31) != __ffs(peers) - 1) return;
    if (key >= a.cap) return;                       // over capacity: reported through n_labels
    const int slot = key & (NSLOT - 1);
    const int old = atomicCAS(&sc.key[slot], -1, key);
    if (old == -1 || old == key) {
        atomicMin(&sc.minx[slot], v.minx); atomicMin(&sc.miny[slot], v.miny);
        atomicMax(&sc.maxx[slot], v.maxx); atomicMax(&sc.maxy[slot], v.maxy);
        atomicAdd(&sc.area[slot], v.area);
        atomicAdd(&sc.sumx[slot], v.sumx); atomicAdd(&sc.sumy[slot], v.sumy);
    } else {
        const int64_t i = img_off + key;
        atomicMin(&a.minx[i], v.minx); atomicMin(&a.miny[i], v.miny);
        atomicMax(&a.maxx[i], v.maxx); atomicMax(&a.maxy[i], v.maxy);
        atomicAdd(&a.area[i], v.area);
        atomicAdd(&a.sumx[i], (unsigned long long)v.sumx); atomicAdd(&a.sumy[i], (unsigned long long)v.sumy);
    }
}

__device__ __forceinline__ Contrib contrib_of(uint32_t px, int c, int r)
{
    Contrib v;
    v.area = __popc(px);
    const int nx1 = ((px >> 1) & 1) + ((px >> 3) & 1);     // pixels in column c+1
    const int ny1 = ((px >> 2) & 1) + ((px >> 3) & 1);     // pixels in row r+1
    v.sumx = (unsigned)(v.area * c + nx1);
    v.sumy = (unsigned)(v.area * r + ny1);
    v.minx = (px & 5u) ? c : c + 1;
    v.maxx = (px & 10u) ? c + 1 : c;
    v.miny = (px & 3u) ? r : r + 1;
    v.maxy = (px & 12u) ? r + 1 : r;
    return v;
}

template <bool WRITE_LABELS>
__global__ void __launch_bounds__(256) ccl_final_kernel(MaskAcc m, CclGeom g, const int32_t *Lall, Plane labels, StatAcc a)
{
    __shared__ SlotCache sc;
    const int tid = threadIdx.x + threadIdx.y * 32;
    if (tid < NSLOT) {
        sc.key[tid] = -1; sc.minx[tid] = 0x7fffffff; sc.miny[tid] = 0x7fffffff; sc.maxx[tid] = -1; sc.maxy[tid] = -1;
        sc.area[tid] = 0; sc.sumx[tid] = 0; sc.sumy[tid] = 0;
    }
    __syncthreads();
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y, img = blockIdx.z;
    const bool inside = bx < g.bw && by < g.bh;
    const int c = 2 * bx, r = 2 * by;
    uint32_t px = 0, valid = 0;
    int label = 0;
    if (inside) {
        px = block_px(m, img, c, r);
        valid = 1u | (c + 1 < m.width ? 2u : 0u);
        if (r + 1 < m.height) valid |= valid << 2;
        if (px) {
            const int32_t *L = Lall + img * g.bper;
            int32_t v = L[by * g.bw + bx];
            if (v >= 0) v = L[v];
            label = -v - 1;
        }
        if (WRITE_LABELS) {
            int32_t *row0 = (int32_t *)(labels.p + img * labels.bs + r * labels.rs) + c;
            const int l0 = (px & 1u) ? label : 0, l1 = (px & 2u) ? label : 0;
            const int l2 = (px & 4u) ? label : 0, l3 = (px & 8u) ? label : 0;
            const bool al8 = ((labels.rs | (int64_t)(uintptr_t)labels.p | labels.bs) & 7) == 0;
            if ((valid & 2u) && al8) *(int2 *)row0 = make_int2(l0, l1);
            else { row0[0] = l0; if (valid & 2u) row0[1] = l1; }
            if (valid & 4u) {
                int32_t *row1 = (int32_t *)((uint8_t *)row0 + labels.rs);
                if ((valid & 8u) && al8) *(int2 *)row1 = make_int2(l2, l3);
                else { row1[0] = l2; if (valid & 8u) row1[1] = l3; }
            }
        }
    }
    const int64_t img_off = (int64_t)img * a.cap;
    // foreground contribution, grouped by label
    const unsigned fgm = __ballot_sync(0xffffffffu, px != 0);
    if (px) {
        const unsigned peers = __match_any_sync(fgm, label);
        accumulate_group(peers, label, contrib_of(px, c, r), sc, a, img_off);
    }
    // background contribution (label 0)
    const uint32_t bgpx = valid & ~px;
    const unsigned bgm = __ballot_sync(0xffffffffu, bgpx != 0);
    if (bgpx) accumulate_group(bgm, 0, contrib_of(bgpx, c, r), sc, a, img_off);
    __syncthreads();
    if (tid < NSLOT && sc.key[tid] >= 0 && sc.area[tid] > 0) {
        const int64_t i = img_off + sc.key[tid];
        atomicMin(&a.minx[i], sc.minx[tid]); atomicMin(&a.miny[i], sc.miny[tid]);
        atomicMax(&a.maxx[i], sc.maxx[tid]); atomicMax(&a.maxy[i], sc.maxy[tid]);
        atomicAdd(&a.area[i], sc.area[tid]);
        atomicAdd(&a.sumx[i], (unsigned long long)sc.sumx[tid]); atomicAdd(&a.sumy[i], (unsigned long long)sc.sumy[tid]);
    }
}

__global__ void stats_finalize_kernel(StatAcc a, const int32_t *n_roots, int batch, int32_t *n_labels, int32_t *stats, double *centroids)
{
    const int img = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = n_roots[img] + 1;
    if (k == 0) n_labels[img] = (n <= a.cap) ? n : -n;
    if (k >= a.cap || k >= n) return;
    const int64_t i = (int64_t)img * a.cap + k;
    int32_t *s = stats + i * 5;
    const int area = a.area[i];
    if (area > 0) {
        s[0] = a.minx[i]; s[1] = a.miny[i]; s[2] = a.maxx[i] - a.minx[i] + 1; s[3] = a.maxy[i] - a.miny[i] + 1; s[4] = area;
        centroids[2 * i] = (double)a.sumx[i] / (double)area;
        centroids[2 * i + 1] = (double)a.sumy[i] / (double)area;
    } else {   // only possible for the background of an all-foreground image; cv2 4.13 reports exactly this row
        s[0] = -1; s[1] = 0x7fffffff; s[2] = 0; s[3] = 0; s[4] = 0;
        centroids[2 * i] = __longlong_as_double(0x7ff8000000000000LL);
        centroids[2 * i + 1] = __longlong_as_double(0x7ff8000000000000LL);
    }
}

// ---- hysteresis kernels ----------------------------------------------------------------------
// mark the root of every block that holds a strong (class 2) pixel: L[root] = -(root) - 2
__global__ void __launch_bounds__(256) hyst_flag_kernel(Plane cls, int width, int height, CclGeom g, int32_t *Lall)
{
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y, img = blockIdx.z;
    if (bx >= g.bw || by >= g.bh) return;
    const int c = 2 * bx, r = 2 * by;
    const uint8_t *row = cls.p + img * cls.bs + r * cls.rs + c;
    bool strong = row[0] == 2;
    if (c + 1 < width) strong |= row[1] == 2;
    if (r + 1 < height) { strong |= row[cls.rs] == 2; if (c + 1 < width) strong |= row[cls.rs + 1] == 2; }
    if (!strong) return;
    int32_t *L = Lall + img * g.bper;
    const int32_t b = by * g.bw + bx;
    const int32_t v = L[b];
    if (v < -1) return;              // this block is a root already marked
    L[v] = -v - 2;                   // v is the (compressed) root index; benign race, same value
}

template <bool OUT_BITS>
__global__ void __launch_bounds__(256) hyst_final_kernel(Plane cls, int width, int height, CclGeom g, const int32_t *Lall,
                                                         Plane out, BitPlane obits, bool or_bits)
{
    // thread per block as elsewhere; for bit output a half-warp of 16 blocks forms one 32-bit word per row
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y, img = blockIdx.z;
    const bool inside = bx < g.bw && by < g.bh;
    const int c = 2 * bx, r = 2 * by;
    uint32_t keep = 0;    // bit0 (r,c), bit1 (r,c+1), bit2 (r+1,c), bit3 (r+1,c+1)
    if (inside) {
        const uint8_t *row = cls.p + img * cls.bs + r * cls.rs + c;
        uint32_t px = row[0] != 0;
        if (c + 1 < width) px |= (row[1] != 0) << 1;
        if (r + 1 < height) { px |= (row[cls.rs] != 0) << 2; if (c + 1 < width) px |= (row[cls.rs + 1] != 0) << 3; }
        if (px) {
            const int32_t *L = Lall + img * g.bper;
            int32_t v = L[by * g.bw + bx];
            if (v >= 0) v = L[v];
            if (v < -1) keep = px;
        }
    }
    if (OUT_BITS) {
        // lanes 0-15 and 16-31 each cover 32 consecutive pixels
        uint32_t w0 = (keep & 3u) << (2 * (threadIdx.x & 15));
        uint32_t w1 = ((keep >> 2) & 3u) << (2 * (threadIdx.x & 15));
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) { w0 |= __shfl_xor_sync(0xffffffffu, w0, d); w1 |= __shfl_xor_sync(0xffffffffu, w1, d); }
        if (inside && (threadIdx.x & 15) == 0) {
            uint32_t *p0 = obits.p + img * obits.bs + (int64_t)r * obits.wpr + (c >> 5);
            *p0 = or_bits ? (*p0 | w0) : w0;
            if (r + 1 < height) { uint32_t *p1 = p0 + obits.wpr; *p1 = or_bits ? (*p1 | w1) : w1; }
        }
    } else if (inside) {
        uint8_t *orow = out.p + img * out.bs + r * out.rs + c;
        orow[0] = (keep & 1u) ? 255 : 0;
        if (c + 1 < width) orow[1] = (keep & 2u) ? 255 : 0;
        if (r + 1 < height) { orow[out.rs] = (keep & 4u) ? 255 : 0; if (c + 1 < width) orow[out.rs + 1] = (keep & 8u) ? 255 : 0; }
    }
}

MaskAcc mask_acc(const CclMask &m)
{
    MaskAcc a;
    if (m.u8) { a.p = (const uint8_t *)m.u8->data; a.rs = m.u8->row_stride; a.bs = m.u8->batch_stride; a.bits = nullptr; a.wpr = 0; a.wbs = 0; }
    else { a.p = nullptr; a.rs = 0; a.bs = 0; a.bits = m.bits.p; a.wpr = m.bits.wpr; a.wbs = m.bits.bs; }
    a.width = m.width; a.height = m.height;
    return a;
}

CclGeom geom_of(int width, int height)
{
    CclGeom g;
    g.bw = (width + 1) / 2; g.bh = (height + 1) / 2; g.nblk = g.bw * g.bh;
    g.bper = (int64_t)align_up((size_t)g.bw * g.bh, 4);
    return g;
}

}  // namespace

size_t ccl_label_scratch_bytes(int width, int height, int batch)
{
    CclGeom g = geom_of(width, height);
    return (size_t)g.bper * batch * sizeof(int32_t);
}

int run_ccl_core(synseg_ctx *ctx, const CclMask &m, int32_t *L, cudaStream_t st)
{
    const MaskAcc a = mask_acc(m);
    const CclGeom g = geom_of(m.width, m.height);
    dim3 block(32, 8), grid(cdiv(g.bw, 32), cdiv(g.bh, 8), m.batch);
    ccl_init_kernel<<<grid, block, 0, st>>>(a, g, L);
    SS_LAUNCH_CHECK(ctx, "ccl_init", st);
    ccl_merge_kernel<<<grid, block, 0, st>>>(a, g, L);
    SS_LAUNCH_CHECK(ctx, "ccl_merge", st);
    const int64_t n = g.bper * m.batch;
    ccl_compress_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(n, L, g.bper, g.nblk);
    SS_LAUNCH_CHECK(ctx, "ccl_compress", st);
    return SYNSEG_OK;
}

int run_ccl_stats(synseg_ctx *ctx, const CclMask &m, const synseg_img *labels, int32_t *n_labels, int32_t *stats,
                  double *centroids, int32_t max_labels, cudaStream_t st)
{
    const CclGeom g = geom_of(m.width, m.height);
    const int batch = m.batch;
    const int nchunks = cdiv(g.bper, CHUNK);
    void *p;
    SS_TRY(arena_alloc(ctx, (size_t)g.bper * batch * 4, &p, st));
    int32_t *L = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, (size_t)nchunks * batch * 4, &p, st));
    int32_t *chunk_cnt = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, (size_t)batch * 4, &p, st));
    int32_t *n_roots = (int32_t *)p;
    StatAcc a;
    a.cap = max_labels;
    const size_t nacc = (size_t)batch * max_labels;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.minx = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.miny = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.maxx = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.maxy = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.area = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 8, &p, st)); a.sumx = (unsigned long long *)p;
    SS_TRY(arena_alloc(ctx, nacc * 8, &p, st)); a.sumy = (unsigned long long *)p;

    SS_TRY(run_ccl_core(ctx, m, L, st));
    ccl_count_kernel<<<dim3(nchunks, batch), 256, 0, st>>>(L, g.bper, g.nblk, nchunks, chunk_cnt);
    SS_LAUNCH_CHECK(ctx, "ccl_count", st);
    ccl_scan_kernel<<<batch, 256, 0, st>>>(nchunks, chunk_cnt, n_roots);
    SS_LAUNCH_CHECK(ctx, "ccl_scan", st);
    ccl_assign_kernel<<<dim3(nchunks, batch), 256, 0, st>>>(L, g.bper, g.nblk, nchunks, chunk_cnt);
    SS_LAUNCH_CHECK(ctx, "ccl_assign", st);
    stats_init_kernel<<<(unsigned)cdiv(nacc, 256), 256, 0, st>>>(a, (int64_t)nacc);
    SS_LAUNCH_CHECK(ctx, "stats_init", st);
    const MaskAcc ma = mask_acc(m);
    dim3 block(32, 8), grid(cdiv(g.bw, 32), cdiv(g.bh, 8), batch);
    if (labels) ccl_final_kernel<true><<<grid, block, 0, st>>>(ma, g, L, plane_of(labels), a);
    else ccl_final_kernel<false><<<grid, block, 0, st>>>(ma, g, L, Plane{nullptr, 0, 0}, a);
    SS_LAUNCH_CHECK(ctx, "ccl_final", st);
    stats_finalize_kernel<<<dim3(cdiv(max_labels, 128), batch), 128, 0, st>>>(a, n_roots, batch, n_labels, stats, centroids);
    SS_LAUNCH_CHECK(ctx, "stats_finalize", st);
    return SYNSEG_OK;
}

size_t ccl_stats_scratch_bytes(int width, int height, int batch, int max_labels)
{
    CclGeom g = geom_of(width, height);
    return (size_t)g.bper * batch * 4 + (size_t)cdiv(g.bper, CHUNK) * batch * 4 + (size_t)batch * max_labels * 36 + 16 * 256 + batch * 4;
}

int run_hysteresis(synseg_ctx *ctx, const synseg_img *cls, const synseg_img *edges_u8, BitPlane edges_bits, bool or_bits,
                   cudaStream_t st)
{
    CclMask m; m.u8 = cls; m.bits = BitPlane{nullptr, 0, 0}; m.width = cls->width; m.height = cls->height; m.batch = cls->batch;
    const CclGeom g = geom_of(m.width, m.height);
    void *p;
    SS_TRY(arena_alloc(ctx, (size_t)g.bper * m.batch * 4, &p, st));
    int32_t *L = (int32_t *)p;
    SS_TRY(run_ccl_core(ctx, m, L, st));
    dim3 block(32, 8), grid(cdiv(g.bw, 32), cdiv(g.bh, 8), m.batch);
    hyst_flag_kernel<<<grid, block, 0, st>>>(plane_of(cls), m.width, m.height, g, L);
    SS_LAUNCH_CHECK(ctx, "hyst_flag", st);
    if (edges_u8) hyst_final_kernel<false><<<grid, block, 0, st>>>(plane_of(cls), m.width, m.height, g, L, plane_of(edges_u8), BitPlane{nullptr, 0, 0}, false);
    else hyst_final_kernel<true><<<grid, block, 0, st>>>(plane_of(cls), m.width, m.height, g, L, Plane{nullptr, 0, 0}, edges_bits, or_bits);
    SS_LAUNCH_CHECK(ctx, "hyst_final", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_ccl_stats(synseg_ctx *ctx, const synseg_img *mask, const synseg_img *labels, int32_t *n_labels,
                                int32_t *stats, double *centroids, int32_t max_labels, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_ccl_stats: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(mask, "mask", 1));
    if (labels) {
        SS_TRY(validate_img(labels, "labels", 4));
        if (!same_shape(mask, labels)) { synseg_set_error("synseg_ccl_stats: labels shape mismatch"); return SYNSEG_E_INVALID; }
        if (((uintptr_t)labels->data | (uintptr_t)labels->row_stride | (uintptr_t)labels->batch_stride) & 3) {
            synseg_set_error("synseg_ccl_stats: labels must be 4-byte aligned"); return SYNSEG_E_INVALID;
        }
    }
    if (!n_labels || !stats || !centroids || max_labels < 1) { synseg_set_error("synseg_ccl_stats: bad result buffers"); return SYNSEG_E_INVALID; }
    if (mask->width > 32766 || mask->height > 32766) { synseg_set_error("synseg_ccl_stats: image larger than 32766"); return SYNSEG_E_INVALID; }
    SS_TRY(arena_ensure(ctx, ccl_stats_scratch_bytes(mask->width, mask->height, mask->batch, max_labels)));
    arena_begin(ctx);
    CclMask m; m.u8 = mask; m.bits = BitPlane{nullptr, 0, 0}; m.width = mask->width; m.height = mask->height; m.batch = mask->batch;
    return run_ccl_stats(ctx, m, labels, n_labels, stats, centroids, max_labels, (cudaStream_t)stream);
}
