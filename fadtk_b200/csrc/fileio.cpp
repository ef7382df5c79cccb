// Batched multi-threaded WAV / .npy I/O behind include/fadtk_b200_io.h (host only; built with g++ into
// libfadtk_io.so).  One worker pool per call, files handed out through an atomic counter; every file is
// independent, its outcome is one status code.  Uses POSIX read/write on whole payloads (no stdio buffering
// between the page cache and the caller's - usually pinned - buffer).
#include "../../include/fadtk_b200_io.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

template <class F>
int for_each_file(int n, int threads, int* status, F&& work) {
    if (n < 0 || (n > 0 && status == nullptr)) return -1;
    if (n == 0) return 0;
    int hw = (int)std::thread::hardware_concurrency();
    if (hw <= 0) hw = 8;
    if (threads <= 0) threads = std::min(hw, 32);
    threads = std::max(1, std::min(threads, n));
    std::atomic<int> next{0};
    auto loop = [&]() {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) status[i] = work(i);
    };
    if (threads == 1) loop();
    else {
        std::vector<std::thread> pool;
        pool.reserve(threads - 1);
        for (int t = 1; t < threads; ++t) pool.emplace_back(loop);
        loop();
        for (auto& th : pool) th.join();
    }
    int bad = 0;
    for (int i = 0; i < n; ++i) bad += status[i] != FAD_IO_OK;
    return bad;
}

struct Fd {
    int fd = -1;
    explicit Fd(int f) : fd(f) {}
    ~Fd() { if (fd >= 0) ::close(fd); }
    Fd(const Fd&) = delete;
    Fd& operator=(const Fd&) = delete;
};

bool read_full(int fd, void* dst, size_t bytes, off_t at) {
    char* p = static_cast<char*>(dst);
    while (bytes > 0) {
        const ssize_t r = ::pread(fd, p, bytes, at);
        if (r <= 0) return false;
        p += r; at += r; bytes -= (size_t)r;
    }
    return true;
}
bool write_full(int fd, const void* src, size_t bytes) {
    const char* p = static_cast<const char*>(src);
    while (bytes > 0) {
        const ssize_t r = ::write(fd, p, bytes);
        if (r <= 0) return false;
        p += r; bytes -= (size_t)r;
    }
    return true;
}

uint32_t le32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t le16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
void put32(unsigned char* p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; p[2] = (v >> 16) & 255; p[3] = (v >> 24) & 255; }
void put16(unsigned char* p, uint16_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; }

// ---- RIFF/WAVE -------------------------------------------------------------------------------------------
struct WavInfo { int sample_rate = 0, channels = 0; long long frames = 0; off_t data_at = 0; };

// Walks the chunk list: "fmt " must describe integer PCM (tag 1, or WAVE_FORMAT_EXTENSIBLE with the PCM
// sub-format) at 16 bits; "data" gives the payload.  A data size of 0 / 0xFFFFFFFF (streamed writers) or one
// running past the end of the file is clamped to what the file holds, as Python's wave / libsndfile do.
int parse_wav(int fd, WavInfo& w) {
    struct stat st;
    if (::fstat(fd, &st) != 0) return FAD_IO_EOPEN;
    const off_t size = st.st_size;
    unsigned char hdr[12];
    if (size < 12 || !read_full(fd, hdr, 12, 0)) return FAD_IO_EFORMAT;
    if (memcmp(hdr, "RIFF", 4) != 0 || memcmp(hdr + 8, "WAVE", 4) != 0) return FAD_IO_EFORMAT;
    off_t at = 12;
    bool have_fmt = false;
    int bits = 0, block = 0;
    while (at + 8 <= size) {
        unsigned char ck[8];
        if (!read_full(fd, ck, 8, at)) return FAD_IO_EFORMAT;
        const uint32_t len = le32(ck + 4);
        const off_t body = at + 8;
        if (memcmp(ck, "fmt ", 4) == 0) {
            unsigned char f[40];
            const size_t take = std::min<size_t>(len, sizeof f);
            if (take < 16 || !read_full(fd, f, take, body)) return FAD_IO_EFORMAT;
            int tag = le16(f);
            w.channels = le16(f + 2);
            w.sample_rate = (int)le32(f + 4);
            block = le16(f + 12);
            bits = le16(f + 14);
            if (tag == 0xFFFE) {                                   // extensible: sub-format GUID starts with the real tag
                if (take < 26) return FAD_IO_EFORMAT;
                tag = le16(f + 24);
            }
            if (tag != 1 || bits != 16) return FAD_IO_EUNSUPPORTED;
            if (w.channels <= 0 || w.sample_rate <= 0 || block != 2 * w.channels) return FAD_IO_EFORMAT;
            have_fmt = true;
        } else if (memcmp(ck, "data", 4) == 0) {
            if (!have_fmt) return FAD_IO_EFORMAT;
            long long bytes = len;
            if (len == 0 || len == 0xFFFFFFFFu || body + (off_t)len > size) bytes = (long long)(size - body);
            w.frames = bytes / block;
            w.data_at = body;
            return FAD_IO_OK;
        }
        at = body + (off_t)len + (len & 1);                        // chunks are word aligned
    }
    return FAD_IO_EFORMAT;
}

// ---- .npy ------------------------------------------------------------------------------------------------
struct NpyInfo { long long rows = 0; int cols = 1, ndim = 0, dtype = 0; off_t data_at = 0; };

// header dict written by numpy.lib.format: {'descr': '<f2', 'fortran_order': False, 'shape': (10, 128), }
int parse_npy(int fd, NpyInfo& a) {
    unsigned char pre[12];
    if (!read_full(fd, pre, 10, 0)) return FAD_IO_EFORMAT;
    if (memcmp(pre, "\x93NUMPY", 6) != 0) return FAD_IO_EFORMAT;
    const int major = pre[6];
    size_t hlen, hat;
    if (major == 1) { hlen = le16(pre + 8); hat = 10; }
    else if (major == 2 || major == 3) {
        if (!read_full(fd, pre, 12, 0)) return FAD_IO_EFORMAT;
        hlen = le32(pre + 8); hat = 12;
    } else return FAD_IO_EFORMAT;
    if (hlen == 0 || hlen > (1u << 20)) return FAD_IO_EFORMAT;
    std::string h(hlen, '\0');
    if (!read_full(fd, &h[0], hlen, (off_t)hat)) return FAD_IO_EFORMAT;
    a.data_at = (off_t)(hat + hlen);
    auto value_of = [&](const char* key) -> size_t {
        const size_t k = h.find(key);
        if (k == std::string::npos) return std::string::npos;
        const size_t c = h.find(':', k);
        return c == std::string::npos ? c : h.find_first_not_of(' ', c + 1);
    };
    const size_t dv = value_of("'descr'"), fv = value_of("'fortran_order'"), sv = value_of("'shape'");
    if (dv == std::string::npos || fv == std::string::npos || sv == std::string::npos) return FAD_IO_EFORMAT;
    if (h.compare(fv, 5, "False") != 0) return FAD_IO_EUNSUPPORTED;
    if (h[dv] != '\'') return FAD_IO_EUNSUPPORTED;                  // structured dtypes are lists
    const std::string descr = h.substr(dv + 1, h.find('\'', dv + 1) - dv - 1);
    if (descr == "<f2" || descr == "=f2") a.dtype = 2;
    else if (descr == "<f4" || descr == "=f4") a.dtype = 4;
    else if (descr == "<f8" || descr == "=f8") a.dtype = 8;
    else return FAD_IO_EUNSUPPORTED;
    if (h[sv] != '(') return FAD_IO_EFORMAT;
    const size_t close = h.find(')', sv);
    if (close == std::string::npos) return FAD_IO_EFORMAT;
    long long dims[3];
    int nd = 0;
    size_t p = sv + 1;
    while (p < close) {
        while (p < close && (h[p] == ' ' || h[p] == ',')) ++p;
        if (p >= close) break;
        char* end = nullptr;
        const long long v = strtoll(h.c_str() + p, &end, 10);
        if (end == h.c_str() + p || v < 0 || nd == 3) return FAD_IO_EFORMAT;
        dims[nd++] = v;
        p = (size_t)(end - h.c_str());
    }
    if (nd == 1) { a.rows = dims[0]; a.cols = 1; }
    else if (nd == 2) { a.rows = dims[0]; a.cols = (int)dims[1]; }
    else return FAD_IO_EUNSUPPORTED;
    a.ndim = nd;
    return FAD_IO_OK;
}

// np.save's format-1.0 header: the dict, padded with spaces so that the payload starts on a 64-byte boundary,
// newline-terminated (numpy/lib/format.py _wrap_header, ARRAY_ALIGN = 64)
std::string npy_header_f16(long long rows, int d) {
    char dict[128];
    snprintf(dict, sizeof dict, "{'descr': '<f2', 'fortran_order': False, 'shape': (%lld, %d), }", rows, d);
    std::string h(dict);
    const size_t unpadded = 10 + h.size() + 1;
    const size_t pad = (64 - unpadded % 64) % 64;
    h.append(pad, ' ');
    h.push_back('\n');
    std::string out("\x93NUMPY\x01\x00", 8);
    out.push_back((char)(h.size() & 255));
    out.push_back((char)((h.size() >> 8) & 255));
    return out + h;
}

}  // namespace

extern "C" {

int fad_io_version(void) { return 1; }

int fad_io_wav_probe(const char* const* paths, int n, int threads, int* sample_rate, int* channels,
                     long long* frames, int* status) {
    if (n > 0 && (!paths || !sample_rate || !channels || !frames)) return -1;
    return for_each_file(n, threads, status, [&](int i) -> int {
        sample_rate[i] = 0; channels[i] = 0; frames[i] = 0;
        Fd f(::open(paths[i], O_RDONLY | O_CLOEXEC));
        if (f.fd < 0) return FAD_IO_EOPEN;
        WavInfo w;
        const int rc = parse_wav(f.fd, w);
        if (rc == FAD_IO_OK) { sample_rate[i] = w.sample_rate; channels[i] = w.channels; frames[i] = w.frames; }
        return rc;
    });
}

int fad_io_wav_read(const char* const* paths, int n, int threads, int16_t* dst, const long long* offsets,
                    const long long* frames, const int* channels, int* status) {
    if (n > 0 && (!paths || !dst || !offsets || !frames || !channels)) return -1;
    return for_each_file(n, threads, status, [&](int i) -> int {
        Fd f(::open(paths[i], O_RDONLY | O_CLOEXEC));
        if (f.fd < 0) return FAD_IO_EOPEN;
        WavInfo w;
        const int rc = parse_wav(f.fd, w);
        if (rc != FAD_IO_OK) return rc;
        if (w.channels != channels[i] || w.frames < frames[i]) return FAD_IO_ESHORT;   // file changed since the probe
        const size_t bytes = (size_t)frames[i] * (size_t)channels[i] * 2;
        return read_full(f.fd, dst + offsets[i], bytes, w.data_at) ? FAD_IO_OK : FAD_IO_ESHORT;
    });
}

int fad_io_wav_write(const char* const* paths, int n, int threads, const int16_t* src, const long long* offsets,
                     const long long* frames, int sample_rate, int* status) {
    if (n > 0 && (!paths || !src || !offsets || !frames)) return -1;
    if (sample_rate <= 0) return -1;
    return for_each_file(n, threads, status, [&](int i) -> int {
        const unsigned long long bytes = (unsigned long long)frames[i] * 2;
        if (frames[i] < 0 || bytes > 0xFFFFFFFFull - 36) return FAD_IO_EUNSUPPORTED;      // RIFF sizes are 32 bit
        unsigned char h[44];
        memcpy(h, "RIFF", 4); put32(h + 4, (uint32_t)(36 + bytes)); memcpy(h + 8, "WAVEfmt ", 8);
        put32(h + 16, 16); put16(h + 20, 1); put16(h + 22, 1); put32(h + 24, (uint32_t)sample_rate);
        put32(h + 28, (uint32_t)sample_rate * 2); put16(h + 32, 2); put16(h + 34, 16);
        memcpy(h + 36, "data", 4); put32(h + 40, (uint32_t)bytes);
        Fd f(::open(paths[i], O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644));
        if (f.fd < 0) return FAD_IO_EOPEN;
        if (!write_full(f.fd, h, sizeof h) || !write_full(f.fd, src + offsets[i], (size_t)bytes)) return FAD_IO_ESHORT;
        return FAD_IO_OK;
    });
}

int fad_io_npy_write_f16(const char* const* paths, int n, int threads, const void* src, const long long* row_offsets,
                         const long long* rows, int d, int* status) {
    if (n > 0 && (!paths || !src || !row_offsets || !rows)) return -1;
    if (d <= 0) return -1;
    const char* base = static_cast<const char*>(src);
    return for_each_file(n, threads, status, [&](int i) -> int {
        if (rows[i] < 0) return FAD_IO_EUNSUPPORTED;
        const std::string h = npy_header_f16(rows[i], d);
        Fd f(::open(paths[i], O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644));
        if (f.fd < 0) return FAD_IO_EOPEN;
        if (!write_full(f.fd, h.data(), h.size()) ||
            !write_full(f.fd, base + (size_t)row_offsets[i] * d * 2, (size_t)rows[i] * d * 2)) return FAD_IO_ESHORT;
        return FAD_IO_OK;
    });
}

int fad_io_npy_probe(const char* const* paths, int n, int threads, long long* rows, int* cols, int* ndim,
                     int* dtype_code, int* status) {
    if (n > 0 && (!paths || !rows || !cols || !ndim || !dtype_code)) return -1;
    return for_each_file(n, threads, status, [&](int i) -> int {
        rows[i] = 0; cols[i] = 0; ndim[i] = 0; dtype_code[i] = 0;
        Fd f(::open(paths[i], O_RDONLY | O_CLOEXEC));
        if (f.fd < 0) return FAD_IO_EOPEN;
        NpyInfo a;
        const int rc = parse_npy(f.fd, a);
        if (rc == FAD_IO_OK) { rows[i] = a.rows; cols[i] = a.cols; ndim[i] = a.ndim; dtype_code[i] = a.dtype; }
        return rc;
    });
}

int fad_io_npy_read_f16(const char* const* paths, int n, int threads, void* dst, const long long* row_offsets,
                        const long long* rows, int d, int* status) {
    if (n > 0 && (!paths || !dst || !row_offsets || !rows)) return -1;
    if (d <= 0) return -1;
    char* base = static_cast<char*>(dst);
    return for_each_file(n, threads, status, [&](int i) -> int {
        Fd f(::open(paths[i], O_RDONLY | O_CLOEXEC));
        if (f.fd < 0) return FAD_IO_EOPEN;
        NpyInfo a;
        const int rc = parse_npy(f.fd, a);
        if (rc != FAD_IO_OK) return rc;
        if (a.dtype != 2 || a.ndim != 2 || a.cols != d) return FAD_IO_EUNSUPPORTED;
        if (a.rows != rows[i]) return FAD_IO_ESHORT;                                  // file changed since the probe
        return read_full(f.fd, base + (size_t)row_offsets[i] * d * 2, (size_t)rows[i] * d * 2, a.data_at) ? FAD_IO_OK : FAD_IO_ESHORT;
    });
}

}  // extern "C"
