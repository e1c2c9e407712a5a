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
#include <map>
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
// buffers) or a negative SGL_E* code.
int64_t sgl_ivsparse_decode(const void* image, uint64_t bytes, int64_t col0, int64_t ncol, int32_t* p, int32_t* i, double* x, int64_t capacity) {
    Image im;
    int rc = parse(image, bytes, &im);
    if (rc != SGL_OK) return rc;
    if (col0 < 0 || ncol < 0 || col0 + ncol > (int64_t)im.ncol || !p) return fail(SGL_EINVAL, "IVSparse decode: column range out of bounds or NULL p");
    const bool fill = i && x;
    int64_t total = 0;
    p[0] = 0;
    std::vector<std::pair<uint32_t, double>> col;
    if (im.level == 3) {
        uint64_t off = 24 + 8ull * im.ncol;
        for (int64_t c = 0; c < col0; ++c) off += rd(im.sizes3 + 8 * c, 8);
        for (int64_t c = 0; c < ncol; ++c) {
            const uint64_t size = rd(im.sizes3 + 8 * (col0 + c), 8);
            if (off + size > bytes) return fail(SGL_EINVAL, "IVSparse decode: column %lld runs past the end of the image", (long long)(col0 + c));
            col.clear();
            const int64_t n = decode_blob3(im, im.base + off, size, fill ? &col : nullptr);
            if (n < 0) return fail(SGL_EINVAL, "IVSparse decode: malformed column %lld", (long long)(col0 + c));
            off += size;
            if (total + n > 0x7fffffffLL) return fail(SGL_EINVAL, "IVSparse decode: more than 2^31 - 1 non-zeros in the range; decode fewer columns per chunk");
            if (fill) {
                if (total + n > capacity) return fail(SGL_EINVAL, "IVSparse decode: capacity %lld too small", (long long)capacity);
                std::sort(col.begin(), col.end(), [](const std::pair<uint32_t, double>& a, const std::pair<uint32_t, double>& b) { return a.first < b.first; });
                for (int64_t t = 0; t < n; ++t) {
                    if (t > 0 && col[(size_t)t].first == col[(size_t)t - 1].first) return fail(SGL_EINVAL, "IVSparse decode: duplicate row in column %lld", (long long)(col0 + c));
                    i[total + t] = (int32_t)col[(size_t)t].first;
                    x[total + t] = col[(size_t)t].second;
                }
            }
            total += n;
            p[c + 1] = (int32_t)total;
        }
        return total;
    }
    // level 2 (VCSC): values / counts / indices live in three separate sections
    const int ib = (int)im.index_bytes, vb = im.value_bytes;
    uint64_t n_val_before = 0, n_idx_before = 0, n_val_all = 0, n_idx_all = 0;
    for (int64_t c = 0; c < (int64_t)im.ncol; ++c) {
        const uint64_t nv = rd(im.vsizes + (uint64_t)ib * c, ib), ni = rd(im.isizes + (uint64_t)ib * c, ib);
        if (c < col0) { n_val_before += nv; n_idx_before += ni; }
        n_val_all += nv;
        n_idx_all += ni;
    }
    const uint64_t head = 24 + 2ull * ib * im.ncol;
    const uint8_t* values = im.base + head;
    const uint8_t* counts = values + n_val_all * (uint64_t)vb;
    const uint8_t* indices = counts + n_val_all * (uint64_t)ib;
    if (head + n_val_all * (uint64_t)(vb + ib) + n_idx_all * (uint64_t)ib > bytes) return fail(SGL_EINVAL, "IVSparse decode: truncated VCSC image");
    const uint8_t* pv = values + n_val_before * (uint64_t)vb;
    const uint8_t* pc = counts + n_val_before * (uint64_t)ib;
    const uint8_t* pi = indices + n_idx_before * (uint64_t)ib;
    for (int64_t c = 0; c < ncol; ++c) {
        const uint64_t nv = rd(im.vsizes + (uint64_t)ib * (col0 + c), ib), ni = rd(im.isizes + (uint64_t)ib * (col0 + c), ib);
        if (total + (int64_t)ni > 0x7fffffffLL) return fail(SGL_EINVAL, "IVSparse decode: more than 2^31 - 1 non-zeros in the range; decode fewer columns per chunk");
        col.clear();
        uint64_t seen = 0;
        for (uint64_t v = 0; v < nv; ++v) {
            const double val = value_of(im, pv + v * (uint64_t)vb);
            const uint64_t cnt = rd(pc + v * (uint64_t)ib, ib);
            if (seen + cnt > ni) return fail(SGL_EINVAL, "IVSparse decode: counts of column %lld exceed its indices", (long long)(col0 + c));
            if (fill)
                for (uint64_t t = 0; t < cnt; ++t) {
                    const uint64_t row = rd(pi + (seen + t) * (uint64_t)ib, ib);
                    if (row >= im.nrow) return fail(SGL_EINVAL, "IVSparse decode: row out of range in column %lld", (long long)(col0 + c));
                    col.emplace_back((uint32_t)row, val);
                }
            seen += cnt;
        }
        if (seen != ni) return fail(SGL_EINVAL, "IVSparse decode: counts of column %lld do not add up", (long long)(col0 + c));
        if (fill) {
            if (total + (int64_t)ni > capacity) return fail(SGL_EINVAL, "IVSparse decode: capacity %lld too small", (long long)capacity);
            std::sort(col.begin(), col.end(), [](const std::pair<uint32_t, double>& a, const std::pair<uint32_t, double>& b) { return a.first < b.first; });
            for (uint64_t t = 0; t < ni; ++t) {
                if (t > 0 && col[t].first == col[t - 1].first) return fail(SGL_EINVAL, "IVSparse decode: duplicate row in column %lld", (long long)(col0 + c));
                i[total + (int64_t)t] = (int32_t)col[t].first;
                x[total + (int64_t)t] = col[t].second;
            }
        }
        pv += nv * (uint64_t)vb;
        pc += nv * (uint64_t)ib;
        pi += ni * (uint64_t)ib;
        total += (int64_t)ni;
        p[c + 1] = (int32_t)total;
    }
    return total;
}

// Encode a chunk list (concatenated by columns, like build_IVCSC / IVCSC::append of src/singlet.cpp:783-835) as the file image
// the reference's IVCSC = IVSparse::SparseMatrix<float, uint64_t, 3, true> (level 3) or VCSC (level 2) type writes: values are
// narrowed to float, the index type recorded in the metadata is 8 bytes. Returns the image size in bytes; writes it when `out`
// is not NULL and `capacity` suffices (call once with out = NULL to size the buffer). Negative SGL_E* code on error.
int64_t sgl_ivsparse_encode(const sgl_csc* chunks, int n_chunks, int level, void* out, uint64_t capacity) {
    if (!chunks || n_chunks < 1) return fail(SGL_EINVAL, "IVSparse encode: empty chunk list");
    if (level != 2 && level != 3) return fail(SGL_EINVAL, "IVSparse encode: level must be 2 (VCSC) or 3 (IVCSC)");
    const int64_t nrow = chunks[0].nrow;
    int64_t ncol = 0, nnz = 0;
    for (int q = 0; q < n_chunks; ++q) {
        if (!chunks[q].p || chunks[q].nrow != nrow) return fail(SGL_EINVAL, "IVSparse encode: bad chunk %d", q);
        ncol += chunks[q].ncol;
        nnz += (int64_t)chunks[q].p[chunks[q].ncol] - chunks[q].p[0];
    }
    if (nrow > 0xffffffffLL || ncol > 0xffffffffLL || nnz > 0xffffffffLL) return fail(SGL_EINVAL, "IVSparse encode: dimensions exceed the format's uint32 metadata");
    auto byte_width = [](uint64_t s) -> int {
        int w = 1;
        while (w < 8 && s > ((1ull << (8 * w)) - 1)) ++w;
        return w;
    };
    // pass 1: per-column dictionaries -> sizes; pass 2: write. The dictionaries are rebuilt in pass 2 (memory stays O(column)).
    uint8_t* o = static_cast<uint8_t*>(out);
    const uint64_t head3 = 24 + 8ull * (uint64_t)ncol, head2 = 24 + 16ull * (uint64_t)ncol;
    auto for_each_column = [&](auto&& fn) {
        int64_t c = 0;
        for (int q = 0; q < n_chunks; ++q)
            for (int64_t cc = 0; cc < chunks[q].ncol; ++cc, ++c) {
                std::map<float, std::vector<uint64_t>> dict;  // value -> rows in column order (ascending)
                for (int64_t t = chunks[q].p[cc]; t < chunks[q].p[cc + 1]; ++t) dict[(float)chunks[q].x[t]].push_back((uint64_t)chunks[q].i[t]);
                fn(c, dict);
            }
    };
    std::vector<uint64_t> s1((size_t)ncol, 0), s2((size_t)ncol, 0);  // level 3: blob bytes; level 2: (distinct values, indices)
    uint64_t total_vals = 0, total_idx = 0, total_blob = 0;
    for_each_column([&](int64_t c, std::map<float, std::vector<uint64_t>>& dict) {
        if (level == 3) {
            uint64_t sz = 0;
            for (auto& kv : dict) {
                uint64_t mx = kv.second[0];
                for (size_t t = 1; t < kv.second.size(); ++t) mx = std::max(mx, kv.second[t] - kv.second[t - 1]);
                sz += 4 + 1 + (uint64_t)byte_width(mx) * (kv.second.size() + 1);
            }
            s1[(size_t)c] = sz;
            total_blob += sz;
        } else {
            s1[(size_t)c] = dict.size();
            uint64_t ni = 0;
            for (auto& kv : dict) ni += kv.second.size();
            s2[(size_t)c] = ni;
            total_vals += dict.size();
            total_idx += ni;
        }
    });
    const uint64_t need = level == 3 ? head3 + total_blob : head2 + total_vals * (4 + 8) + total_idx * 8;
    if (!o) return (int64_t)need;
    if (capacity < need) return fail(SGL_EINVAL, "IVSparse encode: capacity %llu < %llu bytes", (unsigned long long)capacity, (unsigned long long)need);
    const uint32_t md[6] = {(uint32_t)level, (uint32_t)nrow, (uint32_t)ncol, (uint32_t)nnz, 4u | (1u << 8) | (1u << 16) | (1u << 24), 8u};
    std::memcpy(o, md, 24);
    if (level == 3) {
        std::memcpy(o + 24, s1.data(), 8 * (size_t)ncol);
        uint8_t* w = o + head3;
        for_each_column([&](int64_t, std::map<float, std::vector<uint64_t>>& dict) {
            for (auto& kv : dict) {
                uint64_t mx = kv.second[0];
                for (size_t t = 1; t < kv.second.size(); ++t) mx = std::max(mx, kv.second[t] - kv.second[t - 1]);
                const int bw = byte_width(mx);
                std::memcpy(w, &kv.first, 4);
                w += 4;
                *w++ = (uint8_t)bw;
                uint64_t prev = 0;
                for (size_t t = 0; t < kv.second.size(); ++t) {
                    const uint64_t e = t == 0 ? kv.second[0] : kv.second[t] - prev;
                    prev = kv.second[t];
                    std::memcpy(w, &e, (size_t)bw);
                    w += bw;
                }
                std::memset(w, 0, (size_t)bw);  // delimiter
                w += bw;
            }
        });
    } else {
        std::memcpy(o + 24, s1.data(), 8 * (size_t)ncol);
        std::memcpy(o + 24 + 8 * (size_t)ncol, s2.data(), 8 * (size_t)ncol);
        uint8_t* wv = o + head2;
        uint8_t* wc = wv + total_vals * 4;
        uint8_t* wi = wc + total_vals * 8;
        for_each_column([&](int64_t, std::map<float, std::vector<uint64_t>>& dict) {
            for (auto& kv : dict) {
                std::memcpy(wv, &kv.first, 4);
                wv += 4;
                const uint64_t cnt = kv.second.size();
                std::memcpy(wc, &cnt, 8);
                wc += 8;
                std::memcpy(wi, kv.second.data(), 8 * kv.second.size());
                wi += 8 * kv.second.size();
            }
        });
    }
    return (int64_t)need;
}

}  // extern "C"
