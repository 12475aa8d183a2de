// Data-set loaders and the experiment-log surface of the C ABI.  Host code only.
//
// f3: the reference reads its data sets with read_data (Code/data_processor.h:31-99): freopen + gets + strtok, one
//     point per line, tokens separated by `delims`, dim = token count of the first line, one row per line, values
//     atof()'d and narrowed to float.  scht_read_text restates that contract with an mmap'd, multi-threaded parser
//     (minutes -> seconds for the 10^8-row configurations) and defines the two cases the reference leaves undefined:
//     a line with fewer tokens than `dim` is zero-filled (the reference copies uninitialised heap there), extra tokens
//     are ignored (the reference overruns its row buffer).  Plus a flat binary format ("SCHTBIN1", mmap-able) and the
//     .fvecs format the SIFT data sets ship in.
// f4: scht_format_tree_info / scht_format_knn_log write the same line-by-line layout as
//     Convex_Tree<T>::save_tree_info (Code/Convex_Tree.h:964-991) and GPU_Convex_Tree<T>::save_KNN_running_time_info
//     (Code/GPU_Convex_Tree.h:973-1005), numbers formatted like my_strcat / my_strcat_longlong / my_strcat_double
//     (Code/basic_functions.h:313-334), so the experiment drivers' logs keep their shape.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/scht.h"
#include "scht_error.h"
#include "scht_host_tree.h"

namespace {

struct Mapping {
    void* p = nullptr;
    size_t len = 0;
    ~Mapping() {
        if (p && len) munmap(p, len);
    }
};

int map_file(const char* path, Mapping& m) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return scht::fail(SCHT_ERR_INVALID_ARG, std::string("cannot open ") + path + ": " + strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) {
        close(fd);
        return scht::fail(SCHT_ERR_INVALID_ARG, std::string("cannot stat ") + path);
    }
    m.len = (size_t)st.st_size;
    if (m.len) {
        m.p = mmap(nullptr, m.len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m.p == MAP_FAILED) {
            m.p = nullptr;
            close(fd);
            return scht::fail(SCHT_ERR_OOM, std::string("mmap failed for ") + path);
        }
    }
    close(fd);
    return SCHT_OK;
}

inline bool is_delim(char c, const char* delims) {
    // like strtok on a gets()-line: the newline was removed by gets, '\r' stays a token character in the reference;
    // a trailing "\r" token would atof() to 0 and overrun the row there, so it is treated as a delimiter here
    if (c == '\r') return true;
    for (const char* d = delims; *d; d++)
        if (c == *d) return true;
    return false;
}

// tokens of [b, e): calls f(token_begin, token_end) for each, returns their number
template <class F>
int for_tokens(const char* b, const char* e, const char* delims, F&& f) {
    int n = 0;
    while (b < e) {
        while (b < e && is_delim(*b, delims)) b++;
        if (b >= e) break;
        const char* t = b;
        while (b < e && !is_delim(*b, delims)) b++;
        f(t, b);
        n++;
    }
    return n;
}

struct BinHeader {
    char magic[8];  // "SCHTBIN1"
    int32_t dim;
    int32_t reserved0;
    int64_t n_points;
    int64_t reserved1;
};
static_assert(sizeof(BinHeader) == 32, "header is 32 bytes: rows start 32-byte aligned");

// ---- log formatting: the reference concatenates into a char[1000] with strcat; same text, bounded here ----
struct Log {
    std::string s;
    void str(const char* t) { s += t; }
    void i(long long v) { s += std::to_string(v); }                   // my_strcat (itoa) / my_strcat_longlong ("%lld")
    void d(double v) {                                                // my_strcat_double ("%f")
        char b[64];
        snprintf(b, sizeof b, "%f", v);
        s += b;
    }
};

int emit(const Log& l, char* buf, size_t cap, size_t* needed) {
    if (needed) *needed = l.s.size() + 1;
    if (!buf || cap == 0) return needed ? SCHT_OK : scht::fail(SCHT_ERR_INVALID_ARG, "NULL buffer");
    if (l.s.size() + 1 > cap) return scht::fail(SCHT_ERR_INVALID_ARG, "log buffer too small: need " + std::to_string(l.s.size() + 1) + " bytes");
    memcpy(buf, l.s.c_str(), l.s.size() + 1);
    return SCHT_OK;
}

}  // namespace

extern "C" {

void scht_free(void* p) { free(p); }

int scht_read_text(const char* path, const char* delims, float** data_out, int64_t* n_points, int32_t* dim) {
    if (!path || !data_out || !n_points || !dim) return scht::fail(SCHT_ERR_INVALID_ARG, "NULL argument");
    if (!delims || !*delims) delims = " ";  // the drivers pass " " (Code/main.cc:146)
    *data_out = nullptr;
    *n_points = 0;
    *dim = 0;
    Mapping m;
    int rc = map_file(path, m);
    if (rc) return rc;
    const char* base = static_cast<const char*>(m.p);
    const char* end = base + m.len;
    // line starts: one row per line, like the gets() loop (Code/data_processor.h:37-41); a last line without '\n' counts
    std::vector<size_t> starts;
    {
        const char* p = base;
        while (p < end) {
            starts.push_back((size_t)(p - base));
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            p = nl ? nl + 1 : end;
        }
    }
    const int64_t n = (int64_t)starts.size();
    if (n == 0) return scht::fail(SCHT_ERR_INVALID_ARG, std::string("empty data file ") + path);
    auto line_end = [&](int64_t i) {
        const char* e = (i + 1 < n) ? base + starts[(size_t)i + 1] - 1 : end;  // without the '\n'
        if (i + 1 >= n && e > base && e[-1] == '\n') e--;
        return e;
    };
    const int d = for_tokens(base + starts[0], line_end(0), delims, [](const char*, const char*) {});  // get_dim, :5-14
    if (d <= 0) return scht::fail(SCHT_ERR_INVALID_ARG, std::string("first line of ") + path + " has no values");
    if ((double)n * d >= 4294967296.0) return scht::fail(SCHT_ERR_INVALID_ARG, "N*D must stay below 2^32 (reference limit, Code/Array.hh:1175-1183)");
    float* data = static_cast<float*>(malloc((size_t)n * d * sizeof(float)));
    if (!data) return scht::fail(SCHT_ERR_OOM, "cannot allocate the data matrix");
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (n < 4096) nt = 1;
    auto work = [&](int64_t r0, int64_t r1) {
        char tok[128];
        for (int64_t r = r0; r < r1; r++) {
            float* row = data + (size_t)r * d;
            int c = 0;
            for_tokens(base + starts[(size_t)r], line_end(r), delims, [&](const char* tb, const char* te) {
                if (c >= d) {  // extra tokens: ignored (undefined in the reference)
                    c++;
                    return;
                }
                size_t len = std::min((size_t)(te - tb), sizeof(tok) - 1);
                memcpy(tok, tb, len);
                tok[len] = 0;
                row[c++] = (float)atof(tok);  // get_data, Code/data_processor.h:16-29
            });
            for (; c < d; c++) row[c] = 0.f;  // short line: zero-filled (uninitialised in the reference)
        }
    };
    std::vector<std::thread> th;
    const int64_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 1; t < nt; t++) {
        int64_t r0 = std::min<int64_t>(n, t * chunk), r1 = std::min<int64_t>(n, (t + 1) * chunk);
        if (r0 < r1) th.emplace_back(work, r0, r1);
    }
    work(0, std::min<int64_t>(n, chunk));
    for (auto& t : th) t.join();
    *data_out = data;
    *n_points = n;
    *dim = d;
    return SCHT_OK;
}

int scht_write_text(const char* path, const float* data, int64_t n_points, int32_t dim) {
    if (!path || (!data && n_points > 0) || n_points < 0 || dim < 1) return scht::fail(SCHT_ERR_INVALID_ARG, "bad argument");
    FILE* f = fopen(path, "w");
    if (!f) return scht::fail(SCHT_ERR_INVALID_ARG, std::string("cannot create ") + path);
    for (int64_t r = 0; r < n_points; r++) {
        for (int j = 0; j < dim; j++) fprintf(f, j ? " %.9g" : "%.9g", (double)data[(size_t)r * dim + j]);  // 9 digits round-trip a float
        fputc('\n', f);
    }
    if (fclose(f) != 0) return scht::fail(SCHT_ERR_INTERNAL, std::string("write failed: ") + path);
    return SCHT_OK;
}

int scht_write_bin(const char* path, const float* data, int64_t n_points, int32_t dim) {
    if (!path || (!data && n_points > 0) || n_points < 0 || dim < 1) return scht::fail(SCHT_ERR_INVALID_ARG, "bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) return scht::fail(SCHT_ERR_INVALID_ARG, std::string("cannot create ") + path);
    BinHeader h{};
    memcpy(h.magic, "SCHTBIN1", 8);
    h.dim = dim;
    h.n_points = n_points;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    const size_t total = (size_t)n_points * dim;
    if (ok && total) ok = fwrite(data, sizeof(float), total, f) == total;
    ok = (fclose(f) == 0) && ok;
    return ok ? SCHT_OK : scht::fail(SCHT_ERR_INTERNAL, std::string("write failed: ") + path);
}

int scht_map_bin(const char* path, const float** data_out, int64_t* n_points, int32_t* dim, void** mapping_out) {
    if (!path || !data_out || !n_points || !dim || !mapping_out) return scht::fail(SCHT_ERR_INVALID_ARG, "NULL argument");
    *mapping_out = nullptr;
    Mapping* m = new Mapping();
    int rc = map_file(path, *m);
    if (rc) {
        delete m;
        return rc;
    }
    const BinHeader* h = static_cast<const BinHeader*>(m->p);
    if (m->len < sizeof(BinHeader) || memcmp(h->magic, "SCHTBIN1", 8) != 0 || h->dim < 1 || h->n_points < 0 ||
        (size_t)h->n_points * (size_t)h->dim * sizeof(float) + sizeof(BinHeader) != m->len) {
        delete m;
        return scht::fail(SCHT_ERR_INVALID_ARG, std::string(path) + " is not a SCHTBIN1 file (or is truncated)");
    }
    *data_out = reinterpret_cast<const float*>(static_cast<const char*>(m->p) + sizeof(BinHeader));
    *n_points = h->n_points;
    *dim = h->dim;
    *mapping_out = m;
    return SCHT_OK;
}

void scht_unmap_bin(void* mapping) { delete static_cast<Mapping*>(mapping); }

int scht_read_fvecs(const char* path, float** data_out, int64_t* n_points, int32_t* dim) {
    if (!path || !data_out || !n_points || !dim) return scht::fail(SCHT_ERR_INVALID_ARG, "NULL argument");
    *data_out = nullptr;
    Mapping m;
    int rc = map_file(path, m);
    if (rc) return rc;
    if (m.len < 4) return scht::fail(SCHT_ERR_INVALID_ARG, std::string("empty fvecs file ") + path);
    int32_t d;
    memcpy(&d, m.p, 4);
    const size_t rec = 4 + (size_t)d * 4;
    if (d < 1 || m.len % rec != 0) return scht::fail(SCHT_ERR_INVALID_ARG, std::string(path) + " is not an fvecs file");
    const int64_t n = (int64_t)(m.len / rec);
    float* data = static_cast<float*>(malloc((size_t)n * d * sizeof(float)));
    if (!data) return scht::fail(SCHT_ERR_OOM, "cannot allocate the data matrix");
    const char* p = static_cast<const char*>(m.p);
    for (int64_t r = 0; r < n; r++) {
        int32_t dr;
        memcpy(&dr, p + (size_t)r * rec, 4);
        if (dr != d) {
            free(data);
            return scht::fail(SCHT_ERR_INVALID_ARG, std::string(path) + ": rows of different dimension");
        }
        memcpy(data + (size_t)r * d, p + (size_t)r * rec + 4, (size_t)d * 4);
    }
    *data_out = data;
    *n_points = n;
    *dim = d;
    return SCHT_OK;
}

int scht_format_tree_info(const scht_host_tree* tree, float leaf_pts_percent, char* buf, size_t cap, size_t* needed) {
    if (!tree) return scht::fail(SCHT_ERR_INVALID_ARG, "tree is NULL");
    const scht_tree_desc& d = tree->desc;
    Log l;  // Code/Convex_Tree.h:964-991, including its stray first fragment
    l.str("\n   data dim=");
    l.str("Tree info: \n");
    l.str("      data size= ");
    l.i(d.n_points);
    l.str(", dim=");
    l.i(d.dim);
    l.str("\n      depth= ");
    l.i(d.depth);
    l.str("\n      nodes num= ");
    l.i(d.n_nodes);
    l.str("\n      leaf nodes num=");
    l.i(d.n_leaves);
    l.str("\n      percent threshold to be a leaf=");
    l.d((double)leaf_pts_percent);
    l.str("\n      number threshold to be a leaf=");
    l.i(d.leaf_threshold);
    l.str("\n      tree building time: ");
    l.str(" ");
    l.d(tree->build_seconds);
    l.str("s\n");
    return emit(l, buf, cap, needed);
}

int scht_format_knn_log(const scht_stats* st, int64_t n_queries, int32_t K, int64_t batch_size, char* buf, size_t cap, size_t* needed) {
    if (!st) return scht::fail(SCHT_ERR_INVALID_ARG, "stats is NULL");
    Log l;  // Code/GPU_Convex_Tree.h:973-1005
    l.str("\n*-----------------------------------------KNN QUERY RESULT ----------------------------------------\n");
    l.str("*     Query points number:");
    l.i(n_queries);
    l.str(", and K=");
    l.i(K);
    l.str("\n*     GPU Alorithm ");
    l.str("running time:");
    l.d(st->ms_total * 1e-3);  // timer_whole_alg
    l.str("\n*          Where:");
    l.str("\n*          (1) Init query points: ");
    l.d(std::max(0.0, st->ms_total - st->ms_device) * 1e-3);  // host side of the call: staging, copies
    l.str("\n*              where Approximate searching: ");
    l.d(0.0);  // the descent runs on the device (k_descend); its distance count is line (4)
    l.str("\n*          (2) Quadratic programming Calls in devices=");
    l.i(st->hyperplane_tests);
    l.str("\n*              where inner productions (alpha'*q) in quadprog =");
    l.i(st->hyperplane_tests);  // one D-dim inner product per (query, node) test: the per-level memo (Code/ConvexNode.h:151-206)
    l.str("\n*          (3) Distance computations in devices=");
    l.i(st->dist_computations);
    l.str("\n*          (4) Distance computations in approximate searching=");
    l.i(st->descent_dists);
    l.str("\n*          (5) Overall distance computations in host and devices=(2)+ (3)+(4)=");
    l.i(st->hyperplane_tests + st->dist_computations + st->descent_dists);
    l.str("\n*          (6) BATCH_SIZE=");
    l.i(batch_size);
    l.str(", batch iteration times=");
    l.i(st->batches);
    l.str("\n*-----------------------------------------KNN QUERY RESULT----------------------------------------");
    return emit(l, buf, cap, needed);
}

}  // extern "C"
