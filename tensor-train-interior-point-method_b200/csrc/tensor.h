// Host-side device-tensor bookkeeping for the native sweep driver (amen_driver.cu): a reference-counted
// device buffer from the stream-ordered allocator plus a strided view (up to 5 axes).  No arithmetic here.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "api_util.h"

namespace ttipm {
namespace drv {

struct DriverError : std::runtime_error {
    int code;
    DriverError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void check_rc(int rc, const char* what) {
    if (rc != 0) throw DriverError(rc, std::string(what) + ": " + ttipm_last_error());
}

// kernel categories of the per-launch profile (ttipm_amen_profile)
enum ProfCat { CAT_MATVEC = 0, CAT_PHI, CAT_RHS, CAT_GEMM, CAT_QR, CAT_SVD, CAT_EWISE, CAT_DENSE, CAT_KRYLOV, CAT_COUNT };

struct ProfRec {
    int cat;
    double work;                  // algorithmic flops (contractions, factorisations) or bytes (memory-bound helpers)
    void* e0;                     // cudaEvent_t pair, created only while profiling
    void* e1;
};

struct Ctx {
    tt_stream_t st = nullptr;
    bool prof = false;            // time every launch with CUDA events (profiling passes only)
    std::vector<ProfRec> recs;
    double* pinned = nullptr;     // host staging for scalar read-backs
    size_t pinned_cap = 0;
    long launches = 0;
    long syncs = 0;
    size_t bytes_live = 0, bytes_peak = 0;
};

// RAII bracket around the launches of one kernel category: a CUDA-event pair on the driver's stream while
// Ctx::prof is set, nothing otherwise
struct ProfScope {
    Ctx& c;
    bool on;
    size_t idx = 0;
    ProfScope(Ctx& ctx, int cat, double work);
    ~ProfScope();
};

void* dev_alloc(Ctx& c, size_t bytes);
void dev_free(Ctx& c, void* p, size_t bytes);
void to_host(Ctx& c, const double* dev, size_t n, double* host);      // synchronises the stream
void from_host(Ctx& c, const double* host, size_t n, double* dev);
void dev_to_dev(Ctx& c, const double* src, size_t n, double* dst);

struct Buf {
    Ctx* c;
    double* p;
    size_t n;
    Buf(Ctx& ctx, size_t count) : c(&ctx), p((double*)dev_alloc(ctx, count * sizeof(double))), n(count) {}
    ~Buf() { dev_free(*c, p, n * sizeof(double)); }
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
};

struct Tensor {
    std::shared_ptr<Buf> buf;
    double* p = nullptr;
    int nd = 0;
    long d[5] = {0, 0, 0, 0, 0};
    long s[5] = {0, 0, 0, 0, 0};

    bool defined() const { return p != nullptr; }
    long numel() const {
        long n = 1;
        for (int i = 0; i < nd; ++i) n *= d[i];
        return n;
    }
    bool contiguous() const {
        long acc = 1;
        for (int i = nd - 1; i >= 0; --i) {
            if (d[i] != 1 && s[i] != acc) return false;
            acc *= d[i];
        }
        return true;
    }
    static Tensor empty(Ctx& c, std::initializer_list<long> dims) {
        Tensor t;
        t.nd = (int)dims.size();
        long n = 1;
        int i = 0;
        for (long v : dims) t.d[i++] = v;
        for (int k = t.nd - 1; k >= 0; --k) {
            t.s[k] = n;
            n *= t.d[k];
        }
        t.buf = std::make_shared<Buf>(c, (size_t)(n > 0 ? n : 1));
        t.p = t.buf->p;
        return t;
    }
    Tensor reshape(std::initializer_list<long> dims) const {
        if (!contiguous()) throw DriverError(90, "reshape of a non-contiguous view");
        Tensor t = *this;
        t.nd = (int)dims.size();
        long n = 1;
        int i = 0;
        for (long v : dims) t.d[i++] = v;
        for (int k = t.nd - 1; k >= 0; --k) {
            t.s[k] = n;
            n *= t.d[k];
        }
        if (n != numel()) throw DriverError(90, "reshape changes the number of elements");
        return t;
    }
    Tensor permute(std::initializer_list<int> ax) const {
        Tensor t = *this;
        int i = 0;
        for (int a : ax) {
            t.d[i] = d[a];
            t.s[i] = s[a];
            ++i;
        }
        return t;
    }
    Tensor t2() const { return permute({1, 0}); }
    // slice [a, b) along axis ax
    Tensor slice(int ax, long a, long b) const {
        Tensor t = *this;
        t.p = p + a * s[ax];
        t.d[ax] = b - a;
        return t;
    }
    // drop axis ax at index j
    Tensor select(int ax, long j) const {
        Tensor t = *this;
        t.p = p + j * s[ax];
        for (int k = ax; k + 1 < nd; ++k) {
            t.d[k] = d[k + 1];
            t.s[k] = s[k + 1];
        }
        t.nd = nd - 1;
        return t;
    }
    Tensor unsqueeze(int ax) const {
        Tensor t = *this;
        for (int k = nd; k > ax; --k) {
            t.d[k] = d[k - 1];
            t.s[k] = s[k - 1];
        }
        t.d[ax] = 1;
        t.s[ax] = 1;
        t.nd = nd + 1;
        return t;
    }
};

typedef std::pair<int, int> Key;
typedef std::map<Key, Tensor> KeyMap;
typedef std::map<int, Tensor> RowMap;

}  // namespace drv
}  // namespace ttipm
