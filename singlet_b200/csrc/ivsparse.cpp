// ivsparse.cpp -- IVSparse wire formats: host-side codec between a dgCMatrix chunk list and the
// value-compressed column formats of the reference's vendored IVSparse library, as written by write_IVCSC /
// save_IVSparse / build_IVCSC2 and read by read_IVSparse / run_nmf_on_sparsematrix_list (src/singlet.cpp:783-995).
//
// File image (inst/include/src/IVCSC/IVCSC_Methods.hpp:72-95, IVCSC_Constructors.hpp:532-610;
// inst/include/src/VCSC/VCSC_Methods.hpp:77-109):
//   uint32 metadata[6] = {compression level, inner dim (rows), outer dim (columns), nnz, value type, index bytes}
//       value type = sizeof(T) | is_floating << 8 | is_signed << 16 | column_major << 24   (IVCSC_Private_Methods.hpp:65-72)
//   level 3 (IVCSC): uint64 byte size of every column, then the column blobs. A blob is a sequence of RUNS, one per distinct
//       value in ascending value order (std::map), each  {T value, uint8 w, first row (w bytes), positive row deltas (w bytes
//       each), w zero bytes as delimiter}, w = byte width of max(first row, largest delta)  (IVCSC_Private_Methods.hpp:128-299;
//       decoded by InnerIterators/IVCSC_Iterator_Methods.hpp:140-238: a zero after the first index of a run ends the run).
//   level 2 (VCSC): index_t-wide counts of distinct values and of indices per column, then per column the distinct values
//       (ascending), their occurrence counts, and the row indices grouped by value  (VCSC_Private_Methods.hpp:124-236).
// The engine itself keeps its own HBM layout (spmm.cuh / spmm_h16.cuh); these formats are decoded on the host into the
// dgCMatrix chunk views that sgl_matrix_upload / sgl_nmf take -- a column RANGE at a time, so an atlas-scale file becomes a
// chunk list without ever holding more than 2^31 non-zeros in one chunk.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/singlet_cuda.h"

namespace sgl {
int fail(int code, const char* fmt, ...);
}
using sgl::fail;

namespace {

struct Image {
    const uint8_t* base = nullptr;
    uint64_t bytes = 0;
    uint32_t level = 0, nrow = 0, ncol = 0, nnz = 0, val_t = 0, index_bytes = 0;
    int value_bytes = 0;
    bool is_float = false, is_signed = false;
    // level 3
    const uint8_t* sizes3 = nullptr;  // uint64[ncol]
    // level 2
    const uint8_t *vsizes = nullptr, *isizes = nullptr;  // index_t[ncol] each
};

inline uint64_t rd(const uint8_t* p, int w) {
    uint64_t v = 0;
    std::memcpy(&v, p, (size_t)w);  // little-endian host (x86-64 / aarch64), as the reference assumes
    return v;
}

int parse(const void* image, uint64_t bytes, Image* im) {
    if (!image || bytes < 24) return fail(SGL_EINVAL, "IVSparse image: shorter than its metadata");
    im->base = static_cast<const uint8_t*>(image);
    im->bytes = bytes;
    uint32_t md[6];
    std::memcpy(md, image, 24);
    im->level = md[0];
    im->nrow = md[1];
    im->ncol = md[2];
    im->nnz = md[3];
    im->val_t = md[4];
    im->index_bytes = md[5];
    im->value_bytes = (int)(md[4] & 0xFF);
    im->is_float = ((md[4] >> 8) & 0xFF) != 0;
    im->is_signed = ((md[4] >> 16) & 0xFF) != 0;
    if (((md[4] >> 24) & 0xFF) != 1) return fail(SGL_EINVAL, "IVSparse image: row-major storage is not supported (the reference writes column-major)");
    if (im->level != 2 && im->level != 3) return fail(SGL_EINVAL, "IVSparse image: compression level %u (only VCSC = 2 and IVCSC = 3 are file formats of the path)", im->level);
    if (!(im->value_bytes == 1 || im->value_bytes == 2 || im->value_bytes == 4 || im->value_bytes == 8) || (im->is_float && im->value_bytes < 4))
        return fail(SGL_EINVAL, "IVSparse image: unsupported value type 0x%x", im->val_t);
    if (im->nrow < 1 || im->nrow > 0x7fffffffu) return fail(SGL_EINVAL, "IVSparse image: bad row count %u", im->nrow);
    if (im->level == 3) {
        if (bytes < 24 + 8ull * im->ncol) return fail(SGL_EINVAL, "IVSparse image: truncated column size table");
        im->sizes3 = im->base + 24;
    } else {
        if (!(im->index_bytes == 1 || im->index_bytes == 2 || im->index_bytes == 4 || im->index_bytes == 8))
            return fail(SGL_EINVAL, "IVSparse image: unsupported index width %u", im->index_bytes);
        if (bytes < 24 + 2ull * im->index_bytes * im->ncol) return fail(SGL_EINVAL, "IVSparse image: truncated size tables");
        im->vsizes = im->base + 24;
        im->isizes = im->vsizes + (uint64_t)im->index_bytes * im->ncol;
    }
    return SGL_OK;
}

inline double value_of(const Image& im, const uint8_t* p) {
    if (im.is_float) {
        if (im.value_bytes == 4) { float f; std::memcpy(&f, p, 4); return (double)f; }
        double d; std::memcpy(&d, p, 8); return d;
    }
    const uint64_t u = rd(p, im.value_bytes);
    if (!im.is_signed) return (double)u;
    switch (im.value_bytes) {
        case 1: return (double)(int8_t)u;
        case 2: return (double)(int16_t)u;
        case 4: return (double)(int32_t)u;
        default: return (double)(int64_t)u;
    }
}

// decode one IVCSC column blob into (row, value) pairs; returns the count or -1 on a malformed blob
int64_t decode_blob3(const Image& im, const uint8_t* p, uint64_t size, std::vector<std::pair<uint32_t, double>>* out) {
    const uint8_t* end = p + size;
    int64_t n = 0;
    const int vb = im.value_bytes;
    while (p < end) {
        if (p + vb + 1 > end) return -1;
        const double v = value_of(im, p);
        p += vb;
        const int w = *p++;
        if (w < 1 || w > 8 || p + w > end) return -1;
        uint64_t row = rd(p, w);  // the first index of a run is absolute (it may be 0)
        p += w;
        for (;;) {
            if (row >= im.nrow) return -1;
            if (out) out->emplace_back((uint32_t)row, v);
            ++n;
            if (p + w > end) return -1;  // every run ends with a delimiter
            const uint64_t delta = rd(p, w);
            p += w;
            if (delta == 0) break;  // delimiter
            row += delta;
        }
    }
    return n;
}

struct ColumnSpan3 { uint64_t offset, size; };

}  // namespace

// ---- column-parallel helpers --------------------------------------------------------------------------------------------
namespace {
int worker_count(int64_t work_items) {
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)(hw == 0 ? 4 : (hw > 16 ? 16 : hw));
    if (const char* ev = getenv("SGL_IVSPARSE_THREADS")) n = atoi(ev) > 0 ? atoi(ev) : n;
    if ((int64_t)n > work_items) n = (int)(work_items > 0 ? work_items : 1);
    return n;
}
template <typename F>
void parallel_ranges(int64_t n_items, int n_threads, F&& fn) {  // fn(thread, begin, end) over contiguous ranges
    if (n_threads <= 1) { fn(0, (int64_t)0, n_items); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t b = n_items * t / n_threads, e = n_items * (t + 1) / n_threads;
        pool.emplace_back([&fn, t, b, e] { fn(t, b, e); });
    }
    for (auto& th : pool) th.join();
}
// (row, value) pairs of one column sorted by row; returns false on a duplicate row
inline bool emit_sorted(std::vector<std::pair<uint32_t, double>>& col, int32_t* i, double* x) {
    std::sort(col.begin(), col.end(), [](const std::pair<uint32_t, double>& a, const std::pair<uint32_t, double>& b) { return a.first < b.first; });
    for (size_t t = 0; t < col.size(); ++t) {
        if (t > 0 && col[t].first == col[t - 1].first) return false;
        i[t] = (int32_t)col[t].first;
        x[t] = col[t].second;
    }
    return true;
}
}  // namespace

extern "C" {

int sgl_ivsparse_info(const void* image, uint64_t bytes, int32_t* level, int64_t* nrow, int64_t* ncol, int64_t* nnz, int32_t* value_bytes) {
    Image im;
    int rc = parse(image, bytes, &im);
    if (rc != SGL_OK) return rc;
    if (level) *level = (int32_t)im.level;
    if (nrow) *nrow = im.nrow;
    if (ncol) *ncol = im.ncol;
    if (nnz) *nnz = im.nnz;
    if (value_bytes) *value_bytes = im.value_bytes;
    return SGL_OK;
}

// Decode the columns [col0, col0 + ncol) into dgCMatrix slots: p (ncol + 1 entries, p[0] = 0), and -- when i and x are not
// NULL -- the row indices (ascending within a column, as a dgCMatrix requires; the file groups them by value) and values of
// at most `capacity` non-zeros. Returns the number of non-zeros of the range (call once with i = x = NULL to size the
// buffers) or a negative SGL_E* code. Two passes, both parallel over column ranges: count (walk the blobs / read the size
// table), prefix sum, then decode every column straight into its final position.
int64_t sgl_ivsparse_decode(const void* image, uint64_t bytes, int64_t col0, int64_t ncol, int32_t* p, int32_t* i, double* x, int64_t capacity) {
    Image im;
    int rc = parse(image, bytes, &im);
    if (rc != SGL_OK) return rc;
    if (col0 < 0 || ncol < 0 || col0 + ncol > (int64_t)im.ncol || !p) return fail(SGL_EINVAL, "IVSparse decode: column range out of bounds or NULL p");
    const bool fill = i && x;
    p[0] = 0;
    if (ncol == 0) return 0;
    const int nt = worker_count(ncol);
    std::vector<int64_t> counts((size_t)ncol, 0);
    std::vector<int> bad((size_t)nt, 0);  // 0 ok, 1 malformed, 2 out-of-range row, 3 duplicate row
    std::vector<int64_t> bad_col((size_t)nt, -1);
    const int ib = (int)im.index_bytes, vb = im.value_bytes;
    // byte offsets of the columns' data
    std::vector<uint64_t> off3, val_before, idx_before;
    uint64_t n_val_all = 0, n_idx_all = 0;
    if (im.level == 3) {
        off3.resize((size_t)ncol + 1);
        uint64_t off = 24 + 8ull * im.ncol;
        for (int64_t c = 0; c < col0; ++c) off += rd(im.sizes3 + 8 * c, 8);
        for (int64_t c = 0; c < ncol; ++c) {
            off3[(size_t)c] = off;
            off += rd(im.sizes3 + 8 * (col0 + c), 8);
        }
        off3[(size_t)ncol] = off;
        if (off > bytes) return fail(SGL_EINVAL, "IVSparse decode: the column range runs past the end of the image");
    } else {
        val_before.resize((size_t)ncol + 1);
        idx_before.resize((size_t)ncol + 1);
        uint64_t nvb = 0, nib = 0;
        for (int64_t c = 0; c < (int64_t)im.ncol; ++c) {
            const uint64_t nv = rd(im.vsizes + (uint64_t)ib * c, ib), ni = rd(im.isizes + (uint64_t)ib * c, ib);
            if (c >= col0 && c < col0 + ncol) {
                val_before[(size_t)(c - col0)] = nvb;
                idx_before[(size_t)(c - col0)] = nib;
                counts[(size_t)(c - col0)] = (int64_t)ni;
            }
            if (c < col0 + ncol) { nvb += nv; nib += ni; }
            n_val_all += nv;
            n_idx_all += ni;
        }
        val_before[(size_t)ncol] = nvb;
        idx_before[(size_t)ncol] = nib;
        const uint64_t head = 24 + 2ull * ib * im.ncol;
        if (head + n_val_all * (uint64_t)(vb + ib) + n_idx_all * (uint64_t)ib > bytes) return fail(SGL_EINVAL, "IVSparse decode: truncated VCSC image");
    }
    if (im.level == 3) {  // pass 1: count by walking the blobs
        parallel_ranges(ncol, nt, [&](int t, int64_t b, int64_t e) {
            for (int64_t c = b; c < e && !bad[(size_t)t]; ++c) {
                const int64_t n = decode_blob3(im, im.base + off3[(size_t)c], off3[(size_t)c + 1] - off3[(size_t)c], nullptr);
                if (n < 0) { bad[(size_t)t] = 1; bad_col[(size_t)t] = col0 + c; break; }
                counts[(size_t)c] = n;
            }
        });
        for (int t = 0; t < nt; ++t)
            if (bad[(size_t)t]) return fail(SGL_EINVAL, "IVSparse decode: malformed column %lld", (long long)bad_col[(size_t)t]);
    }
    int64_t total = 0;
    for (int64_t c = 0; c < ncol; ++c) {
        total += counts[(size_t)c];
        if (total > 0x7fffffffLL) return fail(SGL_EINVAL, "IVSparse decode: more than 2^31 - 1 non-zeros in the range; decode fewer columns per chunk");
        p[c + 1] = (int32_t)total;
    }
    if (!fill) return total;
    if (total > capacity) return fail(SGL_EINVAL, "IVSparse decode: capacity %lld too small for %lld non-zeros", (long long)capacity, (long long)total);
    const uint64_t head2 = 24 + 2ull * ib * im.ncol;
    const uint8_t* values = im.base + head2;
    const uint8_t* cnts = values + n_val_all * (uint64_t)vb;
    const uint8_t* indices = cnts + n_val_all * (uint64_t)ib;
    parallel_ranges(ncol, nt, [&](int t, int64_t b, int64_t e) {
        std::vector<std::pair<uint32_t, double>> col;
        for (int64_t c = b; c < e && !bad[(size_t)t]; ++c) {
            col.clear();
            if (im.level == 3) {
                decode_blob3(im, im.base + off3[(size_t)c], off3[(size_t)c + 1] - off3[(size_t)c], &col);
            } else {
                const uint64_t nv = val_before[(size_t)c + 1] - val_before[(size_t)c], ni = idx_before[(size_t)c + 1] - idx_before[(size_t)c];
                const uint8_t* pv = values + val_before[(size_t)c] * (uint64_t)vb;
                const uint8_t* pc = cnts + val_before[(size_t)c] * (uint64_t)ib;
                const uint8_t* pi = indices + idx_before[(size_t)c] * (uint64_t)ib;
                uint64_t seen = 0;
                for (uint64_t v = 0; v < nv && !bad[(size_t)t]; ++v) {
                    const double val = value_of(im, pv + v * (uint64_t)vb);
                    const uint64_t cnt = rd(pc + v * (uint64_t)ib, ib);
                    if (seen + cnt > ni) { bad[(size_t)t] = 1; break; }
                    for (uint64_t q = 0; q < cnt; ++q) {
                        const uint64_t row = rd(pi + (seen + q) * (uint64_t)ib, ib);
                        if (row >= im.nrow) { bad[(size_t)t] = 2; break; }
                        col.emplace_back((uint32_t)row, val);
                    }
                    seen += cnt;
                }
                if (!bad[(size_t)t] && seen != ni) bad[(size_t)t] = 1;
            }
            if (!bad[(size_t)t] && !emit_sorted(col, i + p[c], x + p[c])) bad[(size_t)t] = 3;
            if (bad[(size_t)t]) bad_col[(size_t)t] = col0 + c;
        }
    });
    for (int t = 0; t < nt; ++t)
        if (bad[(size_t)t])
            return fail(SGL_EINVAL, "IVSparse decode: %s in column %lld", bad[(size_t)t] == 3 ? "duplicate row" : (bad[(size_t)t] == 2 ? "row out of range" : "malformed data"),
                        (long long)bad_col[(size_t)t]);
    return total;
}

// Encode a chunk list (concatenated by columns, like build_IVCSC / IVCSC::append of src/singlet.cpp:783-835) as the file image
// the reference's IVCSC = IVSparse::SparseMatrix<float, uint64_t, 3, true> (level 3) or VCSC (level 2) type writes: values are
// narrowed to float, the index type recorded in the metadata is 8 bytes. Returns the image size in bytes; writes it when `out`
// is not NULL and `capacity` suffices (call once with out = NULL to size the buffer: the image built for the size query is
// kept per thread and handed out by the next call with the same arguments). Negative SGL_E* code on error.
// A column is grouped by value with a stable sort of its (value, row) pairs -- ascending values like the reference's std::map,
// rows ascending inside a value -- and the columns are encoded in parallel, each thread into its own buffer.
int64_t sgl_ivsparse_encode(const sgl_csc* chunks, int n_chunks, int level, void* out, uint64_t capacity) {
    if (!chunks || n_chunks < 1) return fail(SGL_EINVAL, "IVSparse encode: empty chunk list");
    if (level != 2 && level != 3) return fail(SGL_EINVAL, "IVSparse encode: level must be 2 (VCSC) or 3 (IVCSC)");
    const int64_t nrow = chunks[0].nrow;
    int64_t ncol = 0, nnz = 0;
    uint64_t key = 1469598103934665603ull ^ (uint64_t)level;
    auto mix = [&key](uint64_t v) { key = (key ^ v) * 1099511628211ull; };
    for (int q = 0; q < n_chunks; ++q) {
        if (!chunks[q].p || chunks[q].nrow != nrow || chunks[q].ncol < 0) return fail(SGL_EINVAL, "IVSparse encode: bad chunk %d", q);
        const int64_t cn = (int64_t)chunks[q].p[chunks[q].ncol] - chunks[q].p[0];
        if (cn > 0 && (!chunks[q].i || !chunks[q].x)) return fail(SGL_EINVAL, "IVSparse encode: chunk %d has no i / x", q);
        ncol += chunks[q].ncol;
        nnz += cn;
        mix((uint64_t)(uintptr_t)chunks[q].p); mix((uint64_t)(uintptr_t)chunks[q].i); mix((uint64_t)(uintptr_t)chunks[q].x);
        mix((uint64_t)chunks[q].ncol); mix((uint64_t)cn);
    }
    if (nrow > 0xffffffffLL || ncol > 0xffffffffLL || nnz > 0xffffffffLL) return fail(SGL_EINVAL, "IVSparse encode: dimensions exceed the format's uint32 metadata");
    static thread_local uint64_t cached_key = 0;
    static thread_local std::vector<uint8_t> cached;
    if (!(cached_key == key && !cached.empty() && out)) {
        // global column -> (chunk, local column)
        std::vector<int> chunk_of((size_t)ncol);
        std::vector<int64_t> local_of((size_t)ncol);
        {
            int64_t c = 0;
            for (int q = 0; q < n_chunks; ++q)
                for (int64_t cc = 0; cc < chunks[q].ncol; ++cc, ++c) { chunk_of[(size_t)c] = q; local_of[(size_t)c] = cc; }
        }
        auto byte_width = [](uint64_t s) -> int {
            int w = 1;
            while (w < 8 && s > ((1ull << (8 * w)) - 1)) ++w;
            return w;
        };
        const int nt = worker_count(ncol);
        // per thread: level 3 -> blobs; level 2 -> values / counts / indices sections; per column: sizes
        std::vector<std::vector<uint8_t>> sec_a((size_t)nt), sec_b((size_t)nt), sec_c((size_t)nt);
        std::vector<uint64_t> s1((size_t)ncol, 0), s2((size_t)ncol, 0);
        parallel_ranges(ncol, nt, [&](int t, int64_t b, int64_t e) {
            std::vector<std::pair<float, uint32_t>> ent;
            std::vector<uint8_t>&A = sec_a[(size_t)t], &B = sec_b[(size_t)t], &Cc = sec_c[(size_t)t];
            auto put = [](std::vector<uint8_t>& v, const void* src, size_t n) {
                const uint8_t* s8 = static_cast<const uint8_t*>(src);
                v.insert(v.end(), s8, s8 + n);
            };
            for (int64_t c = b; c < e; ++c) {
                const sgl_csc& ch = chunks[chunk_of[(size_t)c]];
                const int64_t cc = local_of[(size_t)c];
                ent.clear();
                for (int64_t q = ch.p[cc]; q < ch.p[cc + 1]; ++q) ent.emplace_back((float)ch.x[q], (uint32_t)ch.i[q]);
                std::stable_sort(ent.begin(), ent.end(), [](const std::pair<float, uint32_t>& u, const std::pair<float, uint32_t>& v) { return u.first < v.first; });
                const size_t blob_start = A.size();
                uint64_t n_runs = 0;
                for (size_t r0 = 0; r0 < ent.size();) {
                    size_t r1 = r0 + 1;
                    while (r1 < ent.size() && !(ent[r0].first < ent[r1].first)) ++r1;  // same key as std::map: neither is less
                    ++n_runs;
                    if (level == 3) {
                        uint64_t mx = ent[r0].second;
                        for (size_t q = r0 + 1; q < r1; ++q) mx = std::max<uint64_t>(mx, (uint64_t)ent[q].second - ent[q - 1].second);
                        const int bw = byte_width(mx);
                        put(A, &ent[r0].first, 4);
                        A.push_back((uint8_t)bw);
                        uint64_t prev = 0;
                        for (size_t q = r0; q < r1; ++q) {
                            const uint64_t d = q == r0 ? (uint64_t)ent[q].second : (uint64_t)ent[q].second - prev;
                            prev = ent[q].second;
                            put(A, &d, (size_t)bw);
                        }
                        const uint64_t zero = 0;
                        put(A, &zero, (size_t)bw);  // delimiter
                    } else {
                        put(A, &ent[r0].first, 4);
                        const uint64_t cnt = r1 - r0;
                        put(B, &cnt, 8);
                        for (size_t q = r0; q < r1; ++q) {
                            const uint64_t row = ent[q].second;
                            put(Cc, &row, 8);
                        }
                    }
                    r0 = r1;
                }
                if (level == 3) s1[(size_t)c] = A.size() - blob_start;
                else { s1[(size_t)c] = n_runs; s2[(size_t)c] = ent.size(); }
            }
        });
        uint64_t tot_a = 0, tot_b = 0, tot_c = 0;
        for (int t = 0; t < nt; ++t) { tot_a += sec_a[(size_t)t].size(); tot_b += sec_b[(size_t)t].size(); tot_c += sec_c[(size_t)t].size(); }
        const uint64_t head = 24 + (level == 3 ? 8ull : 16ull) * (uint64_t)ncol;
        cached.assign((size_t)(head + tot_a + tot_b + tot_c), 0);
        const uint32_t md[6] = {(uint32_t)level, (uint32_t)nrow, (uint32_t)ncol, (uint32_t)nnz, 4u | (1u << 8) | (1u << 16) | (1u << 24), 8u};
        std::memcpy(cached.data(), md, 24);
        std::memcpy(cached.data() + 24, s1.data(), 8 * (size_t)ncol);
        if (level == 2) std::memcpy(cached.data() + 24 + 8 * (size_t)ncol, s2.data(), 8 * (size_t)ncol);
        uint8_t *wa = cached.data() + head, *wb = wa + tot_a, *wc = wb + tot_b;
        for (int t = 0; t < nt; ++t) {
            if (!sec_a[(size_t)t].empty()) std::memcpy(wa, sec_a[(size_t)t].data(), sec_a[(size_t)t].size());
            if (!sec_b[(size_t)t].empty()) std::memcpy(wb, sec_b[(size_t)t].data(), sec_b[(size_t)t].size());
            if (!sec_c[(size_t)t].empty()) std::memcpy(wc, sec_c[(size_t)t].data(), sec_c[(size_t)t].size());
            wa += sec_a[(size_t)t].size(); wb += sec_b[(size_t)t].size(); wc += sec_c[(size_t)t].size();
        }
        cached_key = key;
    }
    const uint64_t need = cached.size();
    if (!out) return (int64_t)need;
    if (capacity < need) return fail(SGL_EINVAL, "IVSparse encode: capacity %llu < %llu bytes", (unsigned long long)capacity, (unsigned long long)need);
    std::memcpy(out, cached.data(), (size_t)need);
    cached.clear();
    cached.shrink_to_fit();
    cached_key = 0;
    return (int64_t)need;
}

}  // extern "C"
