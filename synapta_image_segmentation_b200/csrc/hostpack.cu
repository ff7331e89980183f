// hostpack.cu -- host-side helper of the crop hand-over (no device code): copies the rows of many host images into one
// packed (pinned) buffer with a few threads.
//
// Reference hand-over: _render_region returns one PIL image per region (pdf_image_segmentation.py:3638-3657) and the
// O:887-1010 drivers take them one at a time; `FeatureHints.hints_batch` packs a list of them into the layout
// `synseg_hints_crops` / `synseg_colors_crops` read (rows padded to 16 bytes).  In Python that copy is one strided numpy
// assignment per crop -- a short memcpy per row driven by the interpreter's inner loop -- and it bounds the whole call
// (2,000 crops, 2.95 GB: ~250 ms of packing against 53 ms of PCIe and 19 ms of kernels).  Here the rows are plain
// memcpy calls spread over threads in pieces of about a megabyte.
#include <atomic>
#include <thread>

#include "internal.cuh"

namespace {

struct Piece { int32_t item, row0, rows; };

}  // namespace

extern "C" SYNSEG_EXPORT int synseg_pack_rows(void *dst_base, const synseg_pack_item *items, int32_t n, int32_t threads)
{
    if (n <= 0) return SYNSEG_OK;
    if (!dst_base || !items) { synseg_set_error("synseg_pack_rows: NULL argument"); return SYNSEG_E_INVALID; }
    std::vector<Piece> pieces;
    for (int32_t i = 0; i < n; ++i) {
        const synseg_pack_item &it = items[i];
        if (!it.src || it.rows < 0 || it.row_bytes < 0 || it.src_row_stride < it.row_bytes || it.dst_row_stride < it.row_bytes) {
            synseg_set_error("synseg_pack_rows: bad item %d", i); return SYNSEG_E_INVALID;
        }
        if (it.rows == 0 || it.row_bytes == 0) continue;
        int64_t per = ((int64_t)1 << 20) / it.row_bytes;          // rows per piece: about a megabyte
        if (per < 1) per = 1;
        for (int64_t r = 0; r < it.rows; r += per)
            pieces.push_back(Piece{i, (int32_t)r, (int32_t)(it.rows - r < per ? it.rows - r : per)});
    }
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (;;) {
            const size_t k = next.fetch_add(1, std::memory_order_relaxed);
            if (k >= pieces.size()) return;
            const Piece &p = pieces[k];
            const synseg_pack_item &it = items[p.item];
            const uint8_t *s = (const uint8_t *)it.src + (int64_t)p.row0 * it.src_row_stride;
            uint8_t *d = (uint8_t *)dst_base + it.dst_offset + (int64_t)p.row0 * it.dst_row_stride;
            if (it.src_row_stride == it.row_bytes && it.dst_row_stride == it.row_bytes) memcpy(d, s, (size_t)it.row_bytes * p.rows);
            else
                for (int32_t r = 0; r < p.rows; ++r) memcpy(d + (int64_t)r * it.dst_row_stride, s + (int64_t)r * it.src_row_stride, (size_t)it.row_bytes);
        }
    };
    int nt = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
    if ((size_t)nt > pieces.size()) nt = (int)pieces.size();
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (std::thread &t : pool) t.join();
    return SYNSEG_OK;
}
