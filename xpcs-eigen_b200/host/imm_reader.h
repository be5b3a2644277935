// imm_reader.h -- sequential reader of IMM detector files for the host `corr`.
// Replaces xpcs::io::Imm (reference io/imm.cpp:59-129, header layout io/imm.h:63-144): a
// 1024-byte header per frame (compression@4, elapsed@128 f64, dlen@152 u32, corecotick@620 i32),
// then for sparse frames `dlen` int32 pixel indices followed by `dlen` int16 values, for dense
// frames `dlen` int16 values.  As in the reference, the compression flag of the FIRST header
// decides the layout of the whole file.  Unlike the reference, payloads are kept as the raw
// int32 / int16 arrays of consecutive frames (what xpcs_push_sparse / xpcs_push_dense take) and
// every read is checked.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace xpcs_host {

struct ImmBatch {
    int frames = 0;
    std::vector<int32_t> idx;        // sparse: concatenated pixel indices
    std::vector<int16_t> val;        // sparse: concatenated values; dense: [frames][pixels]
    std::vector<int64_t> offsets;    // sparse: frames + 1 event offsets
    std::vector<double> clock, ticks;
    void clear()
    {
        frames = 0;
        idx.clear();
        val.clear();
        offsets.assign(1, 0);
        clock.clear();
        ticks.clear();
    }
};

class ImmReader {
public:
    explicit ImmReader(const std::string &path) : path_(path)
    {
        f_ = fopen(path.c_str(), "rb");
        if (!f_) throw std::runtime_error("cannot open IMM file " + path);
        unsigned char h[1024];
        if (fread(h, 1024, 1, f_) != 1) throw std::runtime_error("IMM file has no frame header: " + path);
        int32_t comp;
        memcpy(&comp, h + 4, 4);
        sparse_ = comp != 0;
        rewind(f_);
    }
    ~ImmReader()
    {
        if (f_) fclose(f_);
    }
    bool sparse() const { return sparse_; }

    // reads `count` frames; dense frames must all have `pixels` values
    void next(int count, ImmBatch &b, int64_t pixels)
    {
        b.clear();
        unsigned char h[1024];
        for (int i = 0; i < count; i++) {
            if (fread(h, 1024, 1, f_) != 1) throw std::runtime_error("IMM file ends before the configured frame range: " + path_);
            uint32_t dlen;
            double elapsed;
            int32_t tick;
            memcpy(&dlen, h + 152, 4);
            memcpy(&elapsed, h + 128, 8);
            memcpy(&tick, h + 620, 4);
            if (sparse_) {
                const size_t at = b.idx.size();
                b.idx.resize(at + dlen);
                b.val.resize(at + dlen);
                if (dlen && (fread(b.idx.data() + at, 4, dlen, f_) != dlen || fread(b.val.data() + at, 2, dlen, f_) != dlen))
                    throw std::runtime_error("IMM frame payload truncated: " + path_);
                b.offsets.push_back((int64_t)b.idx.size());
            } else {
                if ((int64_t)dlen != pixels) throw std::runtime_error("dense IMM frame size differs from the detector size");
                const size_t at = b.val.size();
                b.val.resize(at + dlen);
                if (fread(b.val.data() + at, 2, dlen, f_) != dlen) throw std::runtime_error("IMM frame payload truncated: " + path_);
            }
            b.clock.push_back(elapsed);
            b.ticks.push_back((double)tick);
            b.frames++;
        }
    }

    void skip(int count)
    {
        unsigned char h[1024];
        for (int i = 0; i < count; i++) {
            if (fread(h, 1024, 1, f_) != 1) throw std::runtime_error("IMM file ends inside the skipped range: " + path_);
            uint32_t dlen;
            memcpy(&dlen, h + 152, 4);
            if (fseek(f_, (long)dlen * (sparse_ ? 6 : 2), SEEK_CUR) != 0) throw std::runtime_error("seek failed in " + path_);
        }
    }

private:
    std::string path_;
    FILE *f_ = nullptr;
    bool sparse_ = false;
};

// A whole frame range of a SPARSE IMM file in two passes over a read-only mapping: pass 1 walks the 1024-byte
// headers (frames to skip, then `nframes` frames: payload position, dlen, elapsed, corecotick), pass 2 copies the
// int32 index and int16 value payloads into two contiguous arrays with a few threads.  Replaces 3 fread calls and two
// vector resizes per frame of ImmReader::next (1.0 s for the 731 MB file of the 1-Mpixel / 100 k-frame configuration)
// by memory-speed copies; the result is exactly what next() delivers.  idx / val are malloc'ed (8 spare elements),
// owned by the caller.
inline void read_sparse_imm_mapped(const std::string &path, int skip_frames, int64_t nframes, int32_t *&idx, int16_t *&val,
                                   std::vector<int64_t> &offsets, std::vector<double> &clock, std::vector<double> &ticks)
{
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open IMM file " + path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 1024) {
        close(fd);
        throw std::runtime_error("IMM file has no frame header: " + path);
    }
    const size_t size = (size_t)st.st_size;
    void *map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) throw std::runtime_error("cannot map IMM file " + path);
    madvise(map, size, MADV_SEQUENTIAL);
    const unsigned char *base = static_cast<const unsigned char *>(map);
    struct Unmap {
        void *p;
        size_t n;
        ~Unmap() { munmap(p, n); }
    } unmap{map, size};
    size_t pos = 0;
    auto dlen_at = [&](size_t p) {
        uint32_t d;
        memcpy(&d, base + p + 152, 4);
        return (size_t)d;
    };
    for (int i = 0; i < skip_frames; i++) {
        if (pos + 1024 > size) throw std::runtime_error("IMM file ends inside the skipped range: " + path);
        pos += 1024 + dlen_at(pos) * 6;
    }
    std::vector<size_t> payload((size_t)nframes);
    offsets.assign((size_t)nframes + 1, 0);
    clock.resize((size_t)nframes);
    ticks.resize((size_t)nframes);
    for (int64_t f = 0; f < nframes; f++) {
        if (pos + 1024 > size) throw std::runtime_error("IMM file ends before the configured frame range: " + path);
        const size_t d = dlen_at(pos);
        double elapsed;
        int32_t tick;
        memcpy(&elapsed, base + pos + 128, 8);
        memcpy(&tick, base + pos + 620, 4);
        if (pos + 1024 + d * 6 > size) throw std::runtime_error("IMM frame payload truncated: " + path);
        payload[(size_t)f] = pos + 1024;
        offsets[(size_t)f + 1] = offsets[(size_t)f] + (int64_t)d;
        clock[(size_t)f] = elapsed;
        ticks[(size_t)f] = (double)tick;
        pos += 1024 + d * 6;
    }
    const int64_t total = offsets[(size_t)nframes];
    idx = static_cast<int32_t *>(malloc(sizeof(int32_t) * ((size_t)total + 8)));
    val = static_cast<int16_t *>(malloc(sizeof(int16_t) * ((size_t)total + 8)));
    if (!idx || !val) {
        free(idx);
        free(val);
        idx = nullptr;
        val = nullptr;
        throw std::runtime_error("out of memory reading " + path);
    }
    memset(idx + total, 0, sizeof(int32_t) * 8);
    memset(val + total, 0, sizeof(int16_t) * 8);
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(8, total / (4 << 20)));
    auto copy_range = [&](int64_t f0, int64_t f1) {
        for (int64_t f = f0; f < f1; f++) {
            const size_t d = (size_t)(offsets[(size_t)f + 1] - offsets[(size_t)f]);
            if (!d) continue;
            memcpy(idx + offsets[(size_t)f], base + payload[(size_t)f], d * 4);
            memcpy(val + offsets[(size_t)f], base + payload[(size_t)f] + d * 4, d * 2);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(copy_range, nframes * t / nthreads, nframes * (t + 1) / nthreads);
    for (auto &t : th) t.join();
}

// The same walk, a chunk of frames at a time (corr --stream_frames): the mapping stays open, a call hands out the
// next `nframes` frames in caller-owned buffers that are reused from chunk to chunk, so the host holds one chunk of a
// file of any length (the online multi-tau does the same on the device, include/xpcs_b200.h: xpcs_stream_*).
class SparseImmStream {
public:
    SparseImmStream(const std::string &path, int skip_frames) : path_(path)
    {
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open IMM file " + path);
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size < 1024) {
            close(fd);
            throw std::runtime_error("IMM file has no frame header: " + path);
        }
        size_ = (size_t)st.st_size;
        map_ = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (map_ == MAP_FAILED) throw std::runtime_error("cannot map IMM file " + path);
        madvise(map_, size_, MADV_SEQUENTIAL);
        for (int i = 0; i < skip_frames; i++) {
            if (pos_ + 1024 > size_) throw std::runtime_error("IMM file ends inside the skipped range: " + path);
            pos_ += 1024 + dlen_at(pos_) * 6;
        }
    }
    ~SparseImmStream()
    {
        if (map_ && map_ != MAP_FAILED) munmap(map_, size_);
    }
    SparseImmStream(const SparseImmStream &) = delete;
    SparseImmStream &operator=(const SparseImmStream &) = delete;

    // events of the next nframes frames (pass 1 of read_sparse_imm_mapped restricted to them)
    int64_t peek_events(int nframes) const
    {
        size_t p = pos_;
        int64_t n = 0;
        for (int f = 0; f < nframes; f++) {
            if (p + 1024 > size_) throw std::runtime_error("IMM file ends before the configured frame range: " + path_);
            const size_t d = dlen_at(p);
            n += (int64_t)d;
            p += 1024 + d * 6;
        }
        return n;
    }
    // idx / val: room for peek_events(nframes) entries; offsets[nframes + 1] from 0; clock / ticks [nframes]
    void next(int nframes, int32_t *idx, int16_t *val, int64_t *offsets, double *clock, double *ticks)
    {
        const unsigned char *base = static_cast<const unsigned char *>(map_);
        offsets[0] = 0;
        for (int f = 0; f < nframes; f++) {
            if (pos_ + 1024 > size_) throw std::runtime_error("IMM file ends before the configured frame range: " + path_);
            const size_t d = dlen_at(pos_);
            if (pos_ + 1024 + d * 6 > size_) throw std::runtime_error("IMM frame payload truncated: " + path_);
            double elapsed;
            int32_t tick;
            memcpy(&elapsed, base + pos_ + 128, 8);
            memcpy(&tick, base + pos_ + 620, 4);
            clock[f] = elapsed;
            ticks[f] = (double)tick;
            if (d) {
                memcpy(idx + offsets[f], base + pos_ + 1024, d * 4);
                memcpy(val + offsets[f], base + pos_ + 1024 + d * 4, d * 2);
            }
            offsets[f + 1] = offsets[f] + (int64_t)d;
            pos_ += 1024 + d * 6;
        }
    }

private:
    size_t dlen_at(size_t p) const
    {
        uint32_t d;
        memcpy(&d, static_cast<const unsigned char *>(map_) + p + 152, 4);
        return (size_t)d;
    }
    std::string path_;
    void *map_ = nullptr;
    size_t size_ = 0, pos_ = 0;
};

}  // namespace xpcs_host
