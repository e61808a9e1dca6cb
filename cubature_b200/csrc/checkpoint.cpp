// checkpoint.cpp -- header handling of the pool checkpoint file (checkpoint.h) and the host-only readers of the C ABI
// (cubature_cuda_checkpoint_info / _regions); the device side (dump and upload of the pool) is in cubature_cuda.cpp.
#include "checkpoint.h"

#include <cstring>
#include <vector>

uint64_t ckpt_mix(uint64_t h, const void* data, size_t bytes) {
    // word-wise multiply-xorshift (not cryptographic: it detects truncation and bit rot, not tampering)
    const unsigned char* p = (const unsigned char*)data;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0x9e3779b97f4a7c15ULL;
        h ^= h >> 29;
    }
    uint64_t tail = 0;
    if (i < bytes) memcpy(&tail, p + i, bytes - i);
    h = (h ^ tail ^ (uint64_t)bytes) * 0xbf58476d1ce4e5b9ULL;
    return h ^ (h >> 31);
}

uint64_t ckpt_integrand_hash(const cub_integrand* integ) {
    uint64_t h = 0x243f6a8885a308d3ULL;
    const int32_t head[3] = {integ->family, integ->dim, integ->fdim};
    h = ckpt_mix(h, head, sizeof head);
    if (!integ->params.empty()) h = ckpt_mix(h, integ->params.data(), integ->params.size() * sizeof(double));
    h = ckpt_mix(h, integ->user_src.data(), integ->user_src.size());
    h = ckpt_mix(h, integ->entry.data(), integ->entry.size());
    return h;
}

void ckpt_seal_header(CkptHeader* h) { h->header_checksum = ckpt_mix(0x13198a2e03707344ULL, h, offsetof(CkptHeader, header_checksum)); }

size_t ckpt_payload_bytes(const CkptHeader& h) {
    return (size_t)h.n_regions * ((size_t)8 * (2 * h.dim + 2 * h.fdim + 1) + 4);
}

int ckpt_read_header(FILE* f, CkptHeader* h, char* msg) {
    if (fread(h, sizeof *h, 1, f) != 1 || memcmp(h->magic, "CUBCKPT1", 8) != 0) {
        cub_set_message(msg, "not a cubature pool checkpoint");
        return CUB_ERR_BAD_ARGS;
    }
    CkptHeader t = *h;
    ckpt_seal_header(&t);
    if (t.header_checksum != h->header_checksum || h->version != 1 || h->dim < 1 || h->dim > CUB_MAX_DIM || h->fdim < 1 || h->n_regions < 1) {
        cub_set_message(msg, "pool checkpoint header is damaged or of an unknown version");
        return CUB_ERR_BAD_ARGS;
    }
    return CUB_OK;
}

extern "C" {

int cubature_cuda_checkpoint_info(const char* path, cub_checkpoint_info_t* info) {
    if (!path || !info) return CUB_ERR_BAD_ARGS;
    FILE* f = fopen(path, "rb");
    if (!f) return CUB_ERR_BAD_ARGS;
    CkptHeader h;
    char msg[256];
    const int rc = ckpt_read_header(f, &h, msg);
    fclose(f);
    if (rc != CUB_OK) return rc;
    memset(info, 0, sizeof *info);
    info->dim = h.dim;
    info->fdim = h.fdim;
    info->family = h.family;
    info->n_regions = h.n_regions;
    info->num_eval = h.num_eval;
    info->iterations = h.iterations;
    for (int d = 0; d < h.dim && d < 12; ++d) {
        info->transform[d] = h.transform[d];
        info->shift[d] = h.shift[d];
    }
    return CUB_OK;
}

int64_t cubature_cuda_checkpoint_regions(const char* path, int64_t first, int64_t count, double* centers, double* halfwidths, double* val,
                                         double* err, int32_t* split_dim) {
    if (!path || first < 0 || count < 0) return CUB_ERR_BAD_ARGS;
    FILE* f = fopen(path, "rb");
    if (!f) return CUB_ERR_BAD_ARGS;
    CkptHeader h;
    char msg[256];
    int rc = ckpt_read_header(f, &h, msg);
    if (rc != CUB_OK) {
        fclose(f);
        return rc;
    }
    const int64_t n = h.n_regions;
    if (first > n) first = n;
    if (first + count > n) count = n - first;
    std::vector<double> row((size_t)count);
    std::vector<int32_t> irow((size_t)count);
    const long base = (long)sizeof(CkptHeader);
    bool ok = true;
    auto read_rows = [&](int64_t row0, int nrows, double* dst, int width) {  // rows row0 .. row0 + nrows of the double matrix -> dst[count][width]
        if (!dst) return;
        for (int r = 0; r < nrows && ok; ++r) {
            if (fseek(f, base + (long)(((row0 + r) * n + first) * 8), SEEK_SET) != 0 || fread(row.data(), 8, (size_t)count, f) != (size_t)count) {
                ok = false;
                break;
            }
            for (int64_t i = 0; i < count; ++i) dst[i * width + r] = row[(size_t)i];
        }
    };
    read_rows(0, h.dim, centers, h.dim);
    read_rows(h.dim, h.dim, halfwidths, h.dim);
    read_rows(2 * h.dim, h.fdim, val, h.fdim);
    read_rows(2 * h.dim + h.fdim, h.fdim, err, h.fdim);
    if (split_dim && ok) {
        const int64_t drows = 2 * h.dim + 2 * h.fdim + 1;
        if (fseek(f, base + (long)(drows * n * 8 + first * 4), SEEK_SET) != 0 || fread(irow.data(), 4, (size_t)count, f) != (size_t)count) ok = false;
        for (int64_t i = 0; i < count && ok; ++i) split_dim[i] = irow[(size_t)i];
    }
    fclose(f);
    return ok ? count : (int64_t)CUB_ERR_BAD_ARGS;
}

}  // extern "C"
