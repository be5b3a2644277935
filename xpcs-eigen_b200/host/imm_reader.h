// imm_reader.h -- sequential reader of IMM detector files for the host `corr`.
// Replaces xpcs::io::Imm (reference io/imm.cpp:59-129, header layout io/imm.h:63-144): a
// 1024-byte header per frame (compression@4, elapsed@128 f64, dlen@152 u32, corecotick@620 i32),
// then for sparse frames `dlen` int32 pixel indices followed by `dlen` int16 values, for dense
// frames `dlen` int16 values.  As in the reference, the compression flag of the FIRST header
// decides the layout of the whole file.  Unlike the reference, payloads are kept as the raw
// int32 / int16 arrays of consecutive frames (what xpcs_push_sparse / xpcs_push_dense take) and
// every read is checked.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
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

}  // namespace xpcs_host
