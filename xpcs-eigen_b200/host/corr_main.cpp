// corr -- the reference's entry point `corr configuration.hdf5 [data.imm]` on the B200 library.
//
// Mirrors main() of the reference (src/xpcs/main.cpp:100-479): same positional argument, same
// flags (--g2out --darkout --frameout= --imm= --inpath= --outpath= --exchange= --entry=, main.cpp:86-98), same
// HDF5 configuration keys (configuration.cpp:80-242), same result datasets written back into
// the configuration file (main.cpp:345-457, corr.cpp:883-923, :1089-1090), same stage names in
// the log ("Loading data", "Computing G2 MultiTau", "Normalizing Data", "Total";
// benchmark.h:56-86).  Everything between the IMM reader and the result writer runs on the GPU
// through the C-ABI of include/xpcs_b200.h; there is no CPU compute path.
// HDF5 I/O is h5lite (the image has no libhdf5).  Inputs: IMM (sparse and dense), with --ufxc the UFXC
// event stream (io/ufxc.cpp), with --rigaku the Rigaku one (io/rigaku.cpp, stride = average = 1) and with
// --hdf5 a frame stack /entry/data/data (io/hdf5.cpp; contiguous uint16 / uint32, no chunk filters).
#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <algorithm>
#include <string>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/xpcs_b200.h"
#include "h5lite.h"
#include "imm_reader.h"

using h5lite::Type;

static void log_info(const char *fmt, ...)
{
    char msg[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof(msg), fmt, ap);
    va_end(ap);
    auto now = std::chrono::system_clock::now();
    std::time_t t = std::chrono::system_clock::to_time_t(now);
    int ms = (int)(std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000);
    char ts[64];
    strftime(ts, sizeof(ts), "%Y-%m-%d %H:%M:%S", localtime(&t));
    printf("[%s.%03d] [console] [info] %s\n", ts, ms, msg);
    fflush(stdout);
}

// xpcs::Benchmark (benchmark.h:56-86): RAII scope timer, "<name> took <t> ms|s|m"
struct Scope {
    std::string name;
    std::chrono::steady_clock::time_point t0;
    explicit Scope(const std::string &n) : name(n), t0(std::chrono::steady_clock::now()) {}
    ~Scope()
    {
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < 1000.0) log_info("%s took %.0f ms", name.c_str(), ms);
        else if (ms < 60000.0) log_info("%s took %.3f s", name.c_str(), ms / 1e3);
        else log_info("%s took %.3f m", name.c_str(), ms / 6e4);
    }
};

struct Flags {
    bool g2out = false, darkout = false, no_compat = false, ufxc = false, rigaku = false, hdf5 = false;
    bool nocompress = false;  // --nocompress: C2T_all/g2_* contiguous instead of one deflate-6 chunk
    std::string imm, inpath, outpath, exchange, entry = "/xpcs", config;
    std::string outfile;  // --outfile=PATH: write configuration + results there and leave the input file untouched
    int device = 0;
    int gpus = 1;  // --gpus N: pixel-shard a sparse multi-tau job over GPUs device .. device+N-1 (NCCL inside the library)
    int frameout = 0;
    int stream_frames = 0;  // --stream_frames K: online multi-tau, K = 2^k frames at a time on the device (xpcs_stream_*)
};


static int parse_flags(int argc, char **argv, Flags &f)
{
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a.size() > 1 && a[0] == '-') {
            while (!a.empty() && a[0] == '-') a.erase(0, 1);
            std::string name = a, val;
            bool has_val = false;
            size_t eq = a.find('=');
            if (eq != std::string::npos) {
                name = a.substr(0, eq);
                val = a.substr(eq + 1);
                has_val = true;
            }
            auto need = [&]() -> std::string {
                if (has_val) return val;
                if (i + 1 < argc) return argv[++i];
                fprintf(stderr, "corr: flag --%s needs a value\n", name.c_str());
                exit(2);
            };
            if (name == "g2out") f.g2out = !has_val || val == "true" || val == "1";
            else if (name == "nog2out") f.g2out = false;
            else if (name == "darkout") f.darkout = !has_val || val == "true" || val == "1";
            else if (name == "imm") f.imm = need();
            else if (name == "inpath") f.inpath = need();
            else if (name == "outpath") f.outpath = need();
            else if (name == "exchange") f.exchange = need();
            else if (name == "entry") f.entry = need();
            else if (name == "outfile") f.outfile = need();
            else if (name == "device") f.device = atoi(need().c_str());
            else if (name == "gpus") f.gpus = std::max(1, atoi(need().c_str()));
            else if (name == "no_compat") f.no_compat = true;
            else if (name == "nocompress") f.nocompress = true;
            else if (name == "frame_threading" || name == "noframe_threading") {
                // Corr::twotime(data, frameThreading) picks between two summation orders of the same C
                // (corr.cpp:562-572); here the contraction has one (tensor-core) path, equal to both within 1e-5
                if (name == "frame_threading" && (!has_val || val == "true" || val == "1"))
                    fprintf(stderr, "corr: --frame_threading selects a CPU threading strategy of the reference; the two-time "
                                    "contraction runs on the tensor cores either way\n");
            }
            else if (name == "frameout") f.frameout = atoi(need().c_str());
            else if (name == "stream_frames") f.stream_frames = atoi(need().c_str());
            else if (name == "ufxc") f.ufxc = !has_val || val == "true" || val == "1";
            else if (name == "rigaku") f.rigaku = !has_val || val == "true" || val == "1";
            else if (name == "hdf5") f.hdf5 = !has_val || val == "true" || val == "1";
            else if (name == "transposed" || name == "notransposed") {
            }  // the reference stores the flag and never uses it (io/hdf5.cpp:62, 121-228)
            else {
                fprintf(stderr, "corr: unknown flag --%s\n", name.c_str());
                return 2;
            }
        } else pos.push_back(a);
    }
    if (pos.empty()) {
        fprintf(stderr, "Please specify a HDF5 metadata file\n");  // main.cpp:103-106
        return 1;
    }
    f.config = pos[0];
    if (pos.size() > 1 && f.imm.empty()) f.imm = pos[1];  // README form: corr config.hdf5 data.imm
    return 0;
}

// Configuration::init (configuration.cpp:80-242) without the singleton
struct Config {
    int xdim = 0, ydim = 0, frame_start_todo = 0, frame_end_todo = 0, dpl = 0, dark_start = 0, dark_end = 0, darks = 0;
    long stride = 1, avg = 1;
    int static_window = 1, normalize_by_framesum = 0, wsize = 0;
    float lld = 0, sigma = 0, norm_factor = 1;
    bool flatfield_enabled = false, twotime = false;
    std::string output_path, imm_path, smoothing_method, smoothing_filter;
    std::vector<int32_t> dqmap, sqmap;
    std::vector<double> flatfield;
    std::vector<int> qphi_bins;
    int frames() const { return (int)((frame_end_todo - frame_start_todo + 1) / (stride * avg)); }  // :594-601
    int real_frames() const { return frame_end_todo - frame_start_todo + 1; }
};

static std::string get_str(const h5lite::File &f, const std::string &p)
{
    const h5lite::Node *n = f.find(p);
    return (n && !n->is_group && n->ds.type == Type::STR) ? n->ds.as_string() : std::string();
}
static double get_num(const h5lite::File &f, const std::string &p, double dflt = 0.0)
{
    const h5lite::Node *n = f.find(p);
    return (n && !n->is_group && n->ds.type != Type::STR && n->ds.count() > 0) ? n->ds.scalar() : dflt;
}

static Config read_config(const h5lite::File &f, const std::string &e)
{
    Config c;
    c.output_path = get_str(f, e + "/output_data");
    c.xdim = (int)get_num(f, "/measurement/instrument/detector/x_dimension");
    c.ydim = (int)get_num(f, "/measurement/instrument/detector/y_dimension");
    c.dqmap = f.dataset(e + "/dqmap").as_i32();
    c.sqmap = f.dataset(e + "/sqmap").as_i32();
    if ((int64_t)c.dqmap.size() != (int64_t)c.xdim * c.ydim || c.sqmap.size() != c.dqmap.size())
        throw h5lite::Error("dqmap/sqmap size does not match x_dimension * y_dimension");
    if (get_str(f, e + "/transposed") == "ENABLED") {  // configuration.cpp:113-133
        std::vector<int32_t> d(c.dqmap.size()), s(c.sqmap.size());
        for (int i = 0; i < c.xdim * c.ydim; i++) {
            const int row = i % c.xdim, col = i / c.xdim;
            d[row * c.ydim + col] = c.dqmap[i];
            s[row * c.ydim + col] = c.sqmap[i];
        }
        c.dqmap.swap(d);
        c.sqmap.swap(s);
        std::swap(c.xdim, c.ydim);
    }
    c.frame_start_todo = (int)get_num(f, e + "/data_begin_todo");
    c.frame_end_todo = (int)get_num(f, e + "/data_end_todo");
    c.dpl = (int)get_num(f, e + "/delays_per_level");
    c.dark_start = (int)get_num(f, e + "/dark_begin_todo");
    c.dark_end = (int)get_num(f, e + "/dark_end_todo");
    c.wsize = (int)get_num(f, e + "/twotime2onetime_window_size");
    c.stride = (long)get_num(f, e + "/stride_frames", 1);
    c.avg = (long)get_num(f, e + "/avg_frames", 1);
    if (c.stride < 1) c.stride = 1;
    if (c.avg < 1) c.avg = 1;
    c.normalize_by_framesum = (int)get_num(f, e + "/normalize_by_framesum");
    if (c.dark_start == c.dark_end || c.dark_end == 0) c.darks = 0;  // configuration.cpp:164-178
    else {
        c.darks = c.dark_end - c.dark_start + 1;
        c.lld = (float)get_num(f, e + "/lld");
        c.sigma = (float)get_num(f, e + "/sigma");
    }
    const float dpx = (float)get_num(f, "/measurement/instrument/detector/x_pixel_size");
    const float dpy = (float)get_num(f, "/measurement/instrument/detector/y_pixel_size");
    const float adu = (float)get_num(f, "/measurement/instrument/detector/adu_per_photon");
    const float preset = (float)get_num(f, "/measurement/instrument/detector/exposure_time");
    const float eff = (float)get_num(f, "/measurement/instrument/detector/efficiency");
    const float dist = (float)get_num(f, "/measurement/instrument/detector/distance");
    const float flux = (float)get_num(f, "/measurement/instrument/source_begin/beam_intensity_transmitted");
    const float thick = (float)get_num(f, "/measurement/sample/thickness");
    float nf = 1.0f;  // configuration.cpp:193-199, float arithmetic in the same order
    nf = nf / eff / adu / preset;
    nf = nf / (dpx / dist * dpy / dist);
    nf /= flux;
    nf /= thick;
    c.norm_factor = nf;
    c.static_window = (int)get_num(f, e + "/static_mean_window_size");
    c.flatfield_enabled = get_str(f, e + "/flatfield_enabled") == "ENABLED";
    if (c.flatfield_enabled) c.flatfield = f.dataset("/measurement/instrument/detector/flatfield").as_f64();
    c.twotime = get_str(f, e + "/analysis_type") == "Twotime";
    if (c.twotime) {
        c.smoothing_method = get_str(f, e + "/smoothing_method");
        c.smoothing_filter = get_str(f, e + "/smoothing_filter");
        const h5lite::Dataset &q = f.dataset(e + "/qphi_bin_to_process");
        std::vector<int64_t> v = q.as_i64();
        const size_t n = q.dims.empty() ? v.size() : (size_t)q.dims[0];  // first dimension = count (:228-231)
        for (size_t i = 0; i < n && i < v.size(); i++) c.qphi_bins.push_back((int)v[i]);
    }
    c.imm_path = get_str(f, e + "/input_file_local");
    return c;
}

static std::string lower(std::string s)
{
    for (char &ch : s) ch = (char)tolower(ch);
    return s;
}

#define CHECK(call)                                                                    \
    do {                                                                               \
        int rc_ = (call);                                                              \
        if (rc_ != XPCS_OK) {                                                          \
            fprintf(stderr, "corr: %s failed (%d): %s\n", #call, rc_, xpcs_last_error(h)); \
            return 3;                                                                  \
        }                                                                              \
    } while (0)

// One whole sparse input held on the host: what xpcs_push_sparse takes.
struct SparseInput {
    std::vector<int32_t> idx;
    std::vector<int16_t> val;
    std::vector<int64_t> offs;  // raw frames + 1
    std::vector<double> clock, ticks;
    int32_t *raw_idx = nullptr;  // set by the mapped IMM reader instead of the vectors (malloc'ed)
    int16_t *raw_val = nullptr;
    SparseInput() = default;
    SparseInput(const SparseInput &) = delete;
    SparseInput &operator=(const SparseInput &) = delete;
    ~SparseInput()
    {
        free(raw_idx);
        free(raw_val);
    }
    const int32_t *idxp() const { return raw_idx ? raw_idx : idx.data(); }
    const int16_t *valp() const { return raw_val ? raw_val : val.data(); }
    int raw_frames() const { return (int)offs.size() - 1; }
};

static int raw_block(const Config &conf)
{
    int block = conf.stride > 1 ? (int)conf.stride : (int)conf.avg;  // main.cpp:258-261
    if (conf.stride > 1 && conf.avg > 1) block = (int)(conf.stride * conf.avg);
    return block;
}

// --ufxc (main.cpp:206-207; io/ufxc.cpp:59-153): a stream of 32-bit event words -- frame counter in bits 31..21
// (11 bits, unwrapped by +-2048 when it jumps by more than 2000; the first word is frame 0), count in bits 16..15,
// column-major pixel in bits 14..0.  Frames come out in file order, a missing frame is an empty frame, the
// reader's SkipFrames does nothing (the frame range always starts at the first frame), clock = ticks = frame number.
static void load_ufxc(const Config &conf, int frames, SparseInput &in)
{
    FILE *fp = fopen(conf.imm_path.c_str(), "rb");
    if (!fp) throw std::runtime_error("cannot open " + conf.imm_path);
    std::vector<uint32_t> words;
    {
        uint32_t buf[4096];
        size_t got;
        while ((got = fread(buf, sizeof(uint32_t), 4096, fp)) > 0) words.insert(words.end(), buf, buf + got);
        fclose(fp);
    }
    const int64_t raw_todo = (int64_t)frames * raw_block(conf);
    std::vector<int64_t> &count = in.offs;
    count.assign((size_t)raw_todo + 1, 0);
    std::vector<int64_t> frame_of(words.size(), -1);
    if (!words.empty()) {
        const long first = (long)(words[0] >> 21);
        long prev = first, wrap = 0;
        for (size_t i = 0; i < words.size(); i++) {
            const long c = (long)(words[i] >> 21);
            if (i > 0) {
                const long diff = c - prev;
                if (diff < -2000) wrap += 2048;
                else if (diff > 2000) wrap -= 2048;
            }
            const long ff = c + wrap - first;
            prev = c;
            if (ff >= 0 && ff < raw_todo) {
                frame_of[i] = ff;
                count[(size_t)ff + 1]++;
            }
        }
    }
    for (int64_t f = 0; f < raw_todo; f++) count[(size_t)f + 1] += count[(size_t)f];
    in.idx.assign((size_t)count[(size_t)raw_todo] + 1, 0);
    in.val.assign((size_t)count[(size_t)raw_todo] + 1, 0);
    {
        std::vector<int64_t> cur(count.begin(), count.end() - 1);
        const uint32_t H = (uint32_t)conf.ydim, W = (uint32_t)conf.xdim;
        for (size_t i = 0; i < words.size(); i++) {
            if (frame_of[i] < 0) continue;
            const uint32_t pix = words[i] & 0x7fffu;
            const int64_t at = cur[(size_t)frame_of[i]]++;  // file order inside a frame
            in.idx[(size_t)at] = (int32_t)((pix % H) * W + pix / H);
            in.val[(size_t)at] = (int16_t)((words[i] >> 15) & 0x3u);
        }
    }
    in.clock.resize((size_t)raw_todo);
    for (int64_t f = 0; f < raw_todo; f++) in.clock[(size_t)f] = (double)f;
    in.ticks = in.clock;
}

// --hdf5 (main.cpp:211-216; io/hdf5.cpp:62-228): /entry/data/data, uint16 or uint32 [frames][.][.] (contiguous or
// chunked with the deflate / shuffle filters -- h5lite inflates them); every non-zero sample of a frame is an event
// at its linear index, the frames before data_begin_todo are skipped, clock = ticks = frame number in the stack.
static void load_hdf5_stack(const Config &conf, int frames, int pixels, SparseInput &in)
{
    const h5lite::File stack = h5lite::File::load(conf.imm_path);
    const h5lite::Dataset &d = stack.dataset("/entry/data/data");
    if (d.dims.size() != 3 || (d.type != Type::U16 && d.type != Type::U32))
        throw std::runtime_error("/entry/data/data must be a 3-d uint16 or uint32 dataset");
    const uint64_t nfr = d.dims[0], per = d.dims[1] * d.dims[2];
    if (per != (uint64_t)pixels) throw std::runtime_error("/entry/data/data: frame size differs from the detector");
    const int64_t raw_todo = (int64_t)frames * raw_block(conf);
    const int64_t first = conf.frame_start_todo > 1 ? conf.frame_start_todo - 1 : 0;  // main.cpp:241-245
    if ((uint64_t)(first + raw_todo) > nfr) throw std::runtime_error("/entry/data/data holds too few frames");
    in.offs.assign(1, 0);
    const uint16_t *p16 = reinterpret_cast<const uint16_t *>(d.data.data());
    const uint32_t *p32 = reinterpret_cast<const uint32_t *>(d.data.data());
    for (int64_t f = first; f < first + raw_todo; f++) {
        for (uint64_t i = 0; i < per; i++) {
            const uint32_t v = d.type == Type::U16 ? p16[(uint64_t)f * per + i] : p32[(uint64_t)f * per + i];
            if (v == 0) continue;
            if (v > 32767u) throw std::runtime_error("/entry/data/data: a sample above 32767 does not fit the int16 event payload");
            in.idx.push_back((int32_t)i);
            in.val.push_back((int16_t)v);
        }
        in.offs.push_back((int64_t)in.idx.size());
        in.clock.push_back((double)f);
    }
    in.ticks = in.clock;
    in.idx.push_back(0);
    in.val.push_back(0);
}

// --rigaku (main.cpp:208-210; io/rigaku.cpp:139-267): 64-bit event words, frame number in bits 63..40, column-major
// pixel in bits 35..16, count in bits 10..0.  Words of frames up to data_begin_todo - 1 are skipped; with
// stride_frames > 1 so are the words of frames that are not a multiple of the stride (:161-162).  An output frame is
// closed when the frame number of a word differs from the previous one (avg_frames == 1) or passes the next boundary
// data_begin_todo - 1 + k * block (avg_frames > 1) (:166-168) -- so frames without events vanish, and a run that
// does not start where the reader expects opens with an empty output frame; reading stops when `frames` output
// frames are closed; masked pixels are dropped after the frame bookkeeping; what is still open at the end of the
// file counts only if it holds more than one pixel (:232); clock = ticks = output frame number.
// The reader sums the words of an output frame per pixel and divides by avg_frames -- the arithmetic of the Filter
// stage over a block of stride * avg raw frames (sparse_filter.cpp:143-172) -- so every output frame is handed to
// the library as such a block, with the words in its first raw frame.
static void load_rigaku(const Config &conf, int frames, int pixels, SparseInput &in)
{
    FILE *fp = fopen(conf.imm_path.c_str(), "rb");
    if (!fp) throw std::runtime_error("cannot open " + conf.imm_path);
    const int block = raw_block(conf);
    const uint64_t stride = (uint64_t)conf.stride, avg = (uint64_t)conf.avg;
    const uint64_t start = (uint64_t)(conf.frame_start_todo > 0 ? conf.frame_start_todo - 1 : 0);
    const uint32_t H = (uint32_t)conf.ydim, W = (uint32_t)conf.xdim;
    std::vector<int32_t> &idx = in.idx;
    std::vector<int16_t> &val = in.val;
    std::vector<int64_t> &offs = in.offs;
    offs.assign(1, 0);
    uint64_t previous = start + 1, next_expected = start + (uint64_t)block;
    int64_t closed = 0;
    size_t open_at = 0;  // first event of the output frame still open
    bool stop = false;
    auto close_frame = [&]() {
        for (int k = 0; k < block; k++) offs.push_back((int64_t)idx.size());
        open_at = idx.size();
        closed++;
    };
    std::vector<uint64_t> buf(1 << 15);
    size_t got;
    while (!stop && (got = fread(buf.data(), sizeof(uint64_t), buf.size(), fp)) > 0) {
        for (size_t i = 0; i < got; i++) {
            const uint64_t wd = buf[i];
            const uint64_t fr = (wd >> 40) & 0xffffffffull;
            if (fr <= start) continue;
            if (closed >= frames) {
                stop = true;
                break;
            }
            if (stride > 1 && fr != 0 && fr % stride != 0) continue;
            if ((avg > 1 && fr > next_expected) || (avg == 1 && fr != previous)) {
                close_frame();
                previous = fr;
                next_expected += (uint64_t)block;
            }
            uint32_t pix = (uint32_t)((wd >> 16) & 0xfffffu);
            pix = (pix % H) * W + pix / H;
            if (pix >= (uint32_t)pixels || conf.dqmap[pix] < 1 || conf.sqmap[pix] < 1) continue;
            idx.push_back((int32_t)pix);
            val.push_back((int16_t)(wd & 0x7ffu));
        }
    }
    fclose(fp);
    if (closed < frames) {  // the open frame: kept if it has more than one distinct pixel
        bool two = false;
        for (size_t k = open_at + 1; k < idx.size() && !two; k++) two = idx[k] != idx[open_at];
        if (two) close_frame();
        else {
            idx.resize(open_at);
            val.resize(open_at);
        }
    } else {  // all frames closed: whatever was collected for the next one is dropped
        idx.resize((size_t)offs.back());
        val.resize((size_t)offs.back());
    }
    const int64_t raw_total = (int64_t)frames * block;
    while ((int64_t)offs.size() - 1 < raw_total) offs.push_back((int64_t)idx.size());
    in.clock.resize((size_t)raw_total);
    for (int64_t f = 0; f < raw_total; f++) in.clock[(size_t)f] = (double)f;
    in.ticks = in.clock;
    idx.push_back(0);
    val.push_back(0);
}

// the whole frame range of a sparse IMM file: headers walked and payloads gathered from a read-only mapping
static void load_imm_sparse(const Config &conf, int frames, SparseInput &in)
{
    const int frame_from = conf.frame_start_todo - 1;  // main.cpp:241-245
    const int64_t raw_todo = (int64_t)frames * raw_block(conf);
    xpcs_host::read_sparse_imm_mapped(conf.imm_path, frame_from > 0 ? frame_from : 0, raw_todo, in.raw_idx, in.raw_val, in.offs,
                                      in.clock, in.ticks);
}

struct FilterSums {
    std::vector<float> pixel_sum, frame_sum, pm_total, pm_partial;
    std::vector<double> clock, ticks;  // [2][raw frames seen]
    int raw_seen = 0;
};

// Multi-GPU multi-tau job (--gpus N): one thread and one handle per GPU, the input cut into N slabs of frames
// balanced by event count; the library redistributes the events to the pixel owners over NVLink and reduces the
// sums (comm.cu), so that every rank ends with the whole-detector results.  Returns 0 or an exit code.
static int run_sharded(const Flags &fl, XpcsParams prm, const SparseInput &in, int n_gpus, FilterSums &sums, int T, int Q,
                       int pixels, std::vector<float> *G2, std::vector<float> *IP, std::vector<float> *IF, std::vector<float> &g2,
                       std::vector<float> &se)
{
    const int raw = in.raw_frames();
    const int64_t E = in.offs[(size_t)raw];
    std::vector<int> cut(n_gpus + 1, raw);
    cut[0] = 0;
    for (int r = 1; r < n_gpus; r++) {
        const int64_t target = E * r / n_gpus;
        int f = (int)(std::lower_bound(in.offs.begin(), in.offs.end(), target) - in.offs.begin());
        cut[r] = std::min(std::max(f, cut[r - 1]), raw);
    }
    unsigned char id[128];
    if (int rc = xpcs_comm_unique_id(id)) {
        fprintf(stderr, "corr: xpcs_comm_unique_id failed (%d): %s\n", rc, xpcs_last_error(nullptr));
        return 3;
    }
    std::vector<xpcs_handle> hs(n_gpus, nullptr);
    std::vector<int> status(n_gpus, 0);
    std::vector<std::string> errors(n_gpus);
    std::vector<std::vector<float>> pG2(n_gpus), pIP(n_gpus), pIF(n_gpus);
    const int F = prm.frames, S_windows = F / prm.static_window;
    XpcsInfo info0;
    memset(&info0, 0, sizeof(info0));
    std::vector<double> stage_ms(3, 0.0);
    auto worker = [&](int r) {
        XpcsParams p = prm;
        p.shard_index = r;
        p.shard_count = n_gpus;
        xpcs_handle h = nullptr;
        auto bad = [&](const char *what, int rc) {
            status[r] = rc;
            errors[r] = std::string(what) + ": " + (h ? xpcs_last_error(h) : xpcs_last_error(nullptr));
        };
        int rc = xpcs_create(&p, fl.device + r, &h);
        if (rc) return bad("xpcs_create", rc);
        hs[r] = h;
        if ((rc = xpcs_comm_init(h, n_gpus, r, id))) return bad("xpcs_comm_init", rc);
        XpcsInfo info;
        xpcs_get_info(h, &info);
        const int S = info.n_static;
        auto t0 = std::chrono::steady_clock::now();
        const int f0 = cut[r], f1 = cut[r + 1];
        rc = xpcs_push_sparse_slab(h, f0, in.idxp(), in.valp(), in.offs.data() + f0, in.clock.data() + f0,
                                   in.ticks.data() + f0, f1 - f0);
        if (rc) return bad("xpcs_push_sparse_slab", rc);
        std::vector<float> ps, fs, pt, pp;
        if (r == 0) {
            sums.pixel_sum.assign((size_t)pixels, 0.f);
            sums.frame_sum.assign(2 * (size_t)F, 0.f);
            sums.pm_total.assign((size_t)std::max(S, 1), 0.f);
            sums.pm_partial.assign((size_t)std::max(S_windows, 1) * std::max(S, 1), 0.f);
            rc = xpcs_finish_ingest(h, sums.pixel_sum.data(), sums.frame_sum.data(), sums.pm_total.data(), sums.pm_partial.data());
        } else {  // the same collectives, results dropped
            ps.assign((size_t)pixels, 0.f);
            fs.assign(2 * (size_t)F, 0.f);
            pt.assign((size_t)std::max(S, 1), 0.f);
            pp.assign((size_t)std::max(S_windows, 1) * std::max(S, 1), 0.f);
            rc = xpcs_finish_ingest(h, ps.data(), fs.data(), pt.data(), pp.data());
        }
        if (rc) return bad("xpcs_finish_ingest", rc);
        auto t1 = std::chrono::steady_clock::now();
        if (G2) {
            pG2[r].resize((size_t)T * pixels);
            pIP[r].resize((size_t)T * pixels);
            pIF[r].resize((size_t)T * pixels);
            rc = xpcs_multitau(h, pG2[r].data(), pIP[r].data(), pIF[r].data());
        } else rc = xpcs_multitau(h, nullptr, nullptr, nullptr);
        if (rc) return bad("xpcs_multitau", rc);
        auto t2 = std::chrono::steady_clock::now();
        std::vector<float> g2r((size_t)T * Q), ser((size_t)T * Q);
        rc = xpcs_normalize(h, r == 0 ? g2.data() : g2r.data(), r == 0 ? se.data() : ser.data());
        if (rc) return bad("xpcs_normalize", rc);
        auto t3 = std::chrono::steady_clock::now();
        if (r == 0) {
            xpcs_get_info(h, &info0);
            stage_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
            stage_ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
            stage_ms[2] = std::chrono::duration<double, std::milli>(t3 - t2).count();
        }
    };
    std::vector<std::thread> th;
    for (int r = 0; r < n_gpus; r++) th.emplace_back(worker, r);
    for (auto &t : th) t.join();
    int bad_rc = 0;
    for (int r = 0; r < n_gpus; r++)
        if (status[r]) {
            fprintf(stderr, "corr: rank %d: %s (%d)\n", r, errors[r].c_str(), status[r]);
            bad_rc = 3;
        }
    if (!bad_rc) {
        log_info("Loading data took %.0f ms", stage_ms[0]);
        log_info("Computing G2 MultiTau took %.0f ms", stage_ms[1]);
        log_info("Normalizing Data took %.0f ms", stage_ms[2]);
        // timestamps: the input's own (every rank only saw its slab's)
        sums.raw_seen = raw;
        sums.clock.assign(2 * (size_t)raw, 0.0);
        sums.ticks.assign(2 * (size_t)raw, 0.0);
        for (int i = 0; i < raw; i++) {
            sums.clock[i] = sums.ticks[i] = i + 1;
            sums.clock[raw + i] = in.clock[(size_t)i];
            sums.ticks[raw + i] = in.ticks[(size_t)i];
        }
        if (G2) {  // every pixel has one owner; the other ranks hold zeros there
            *G2 = std::move(pG2[0]);
            *IP = std::move(pIP[0]);
            *IF = std::move(pIF[0]);
            for (int r = 1; r < n_gpus; r++) {
                const size_t n = (size_t)T * pixels;
                for (size_t i = 0; i < n; i++) {
                    (*G2)[i] += pG2[r][i];
                    (*IP)[i] += pIP[r][i];
                    (*IF)[i] += pIF[r][i];
                }
                std::vector<float>().swap(pG2[r]);
                std::vector<float>().swap(pIP[r]);
                std::vector<float>().swap(pIF[r]);
            }
        }
    }
    for (xpcs_handle h : hs) xpcs_destroy(h);
    return bad_rc;
}

int main(int argc, char **argv)
{
    Flags fl;
    if (int rc = parse_flags(argc, argv, fl)) return rc;
    Scope total("Total");
    log_info("H5 metadata path %s", fl.entry.c_str());
    h5lite::File file;
    Config conf;
    try {
        Scope sc("Configuration Total");
        file = h5lite::File::load(fl.config);
        conf = read_config(file, fl.entry);
    } catch (const std::exception &e) {
        fprintf(stderr, "corr: %s\n", e.what());
        return 1;
    }
    // The results go back into the configuration file (h5_result.cpp:67-103).  h5lite rewrites the file whole:
    // attributes and comments are carried over verbatim, everything else it read is written back with the same
    // values; content it cannot reproduce is reported, and the input is then kept as <file>.orig.
    const std::string result_file = fl.outfile.empty() ? fl.config : fl.outfile;
    for (const std::string &n : file.notes) log_info("note: %s", n.c_str());
    if (!file.lossy.empty()) {
        for (const std::string &n : file.lossy) fprintf(stderr, "corr: warning: %s\n", n.c_str());
        if (fl.outfile.empty()) {
            const std::string bak = fl.config + ".orig";
            struct stat sb;
            if (stat(bak.c_str(), &sb) != 0) {
                FILE *src = fopen(fl.config.c_str(), "rb"), *dst = fopen(bak.c_str(), "wb");
                bool ok = src && dst;
                char buf[1 << 16];
                size_t got;
                while (ok && (got = fread(buf, 1, sizeof(buf), src)) > 0) ok = fwrite(buf, 1, got, dst) == got;
                if (src) fclose(src);
                if (dst) fclose(dst);
                if (!ok) {
                    fprintf(stderr, "corr: cannot keep a backup of %s; use --outfile to write the results elsewhere\n", fl.config.c_str());
                    return 1;
                }
            }
            fprintf(stderr, "corr: warning: the input file holds content that cannot be carried over; original kept as %s\n", bak.c_str());
        }
    }
    if (!fl.imm.empty()) conf.imm_path = fl.imm;
    if (!fl.inpath.empty() && !fl.outpath.empty()) {  // main.cpp:127-140
        size_t pos = conf.imm_path.find(fl.inpath);
        if (pos != std::string::npos) conf.imm_path.replace(pos, fl.inpath.size(), fl.outpath);
    }
    if (!fl.exchange.empty()) conf.output_path = fl.exchange;
    if (conf.output_path.empty()) conf.output_path = "/exchange";
    log_info("Processing IMM file at path %s..", conf.imm_path.c_str());
    struct stat st;
    if (stat(conf.imm_path.c_str(), &st) == 0) log_info("File size %.5g Mbytes", (double)st.st_size / (1024.0 * 1024.0));

    const int frames = conf.frames();
    const int real_frames = conf.real_frames();
    const int pixels = conf.xdim * conf.ydim;
    log_info("Data frames=%d stride=%ld average=%ld", frames, conf.stride, conf.avg);
    if (frames <= 0 || pixels <= 0 || conf.dpl <= 0) {
        fprintf(stderr, "corr: configuration gives no work (frames %d, pixels %d, delays_per_level %d)\n", frames, pixels, conf.dpl);
        return 1;
    }
    int method = 0;
    if (conf.twotime) {
        const std::string m = lower(conf.smoothing_method);
        if (m == "symmetric") method = 1;
        else if (m == "staticmap") method = 2;
        else {
            fprintf(stderr, "Smoothing method is not valid\n");  // main.cpp:160-163
            return 1;
        }
        if (conf.wsize <= 0) {
            fprintf(stderr, "corr: %s/twotime2onetime_window_size must be > 0\n", fl.entry.c_str());
            return 1;
        }
    }

    XpcsParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.struct_size = (int32_t)sizeof(prm);
    prm.width = conf.xdim;
    prm.height = conf.ydim;
    prm.frames = frames;
    prm.delays_per_level = conf.dpl;
    prm.stride_frames = (int32_t)conf.stride;
    prm.avg_frames = (int32_t)conf.avg;
    prm.static_window = conf.static_window > 0 ? conf.static_window : 1;
    prm.normalize_by_framesum = conf.normalize_by_framesum;
    prm.compat_flags = fl.no_compat ? 0u : XPCS_COMPAT_STALE_TAIL;
    if (fl.stream_frames > 0) {
        // Online multi-tau (include/xpcs_b200.h, xpcs_stream_*): the device holds one chunk of frames and a per-pixel
        // state, so the job is no longer bounded by device memory.  The reference's dropped G2 pairs
        // depend on the complete row of a pixel and cannot be reproduced chunk by chunk: the stream gives the exact sums.
        if (conf.twotime || fl.frameout > 0 || fl.gpus > 1) {
            fprintf(stderr, "corr: --stream_frames is for multi-tau jobs on one GPU without --frameout\n");
            return 2;
        }
        if (!fl.no_compat) log_info("--stream_frames %d: exact multi-tau sums (the reference's stale-tail pair losses need complete rows)", fl.stream_frames);
        prm.compat_flags = 0u;
    }
    if (fl.rigaku) prm.compat_flags |= XPCS_COMPAT_LATE_WINDOW;  // the reader counts the static windows its own way (io/rigaku.cpp:190-193)
    prm.lld = conf.lld;
    prm.sigma = conf.sigma;
    prm.dqmap = conf.dqmap.data();
    prm.sqmap = conf.sqmap.data();
    prm.flatfield = conf.flatfield_enabled ? conf.flatfield.data() : nullptr;
    prm.shard_index = 0;
    prm.shard_count = 1;
    const std::string out = conf.output_path;
    const uint64_t uy = (uint64_t)conf.ydim, ux = (uint64_t)conf.xdim;

    // ---- the sharded multi-tau job: sparse inputs only (a dense stack would cross PCIe once per GPU); two-time
    // jobs spread their dynamic partitions over the GPUs further down
    bool sharded = fl.gpus > 1 && !conf.twotime && fl.frameout <= 0;
    std::unique_ptr<xpcs_host::ImmReader> imm_reader;
    if (!fl.ufxc && !fl.hdf5 && !fl.rigaku) {
        try {
            imm_reader.reset(new xpcs_host::ImmReader(conf.imm_path));
        } catch (const std::exception &e) {
            fprintf(stderr, "corr: %s\n", e.what());
            return 1;
        }
        if (!imm_reader->sparse()) sharded = false;
    }
    if (fl.gpus > 1 && !sharded && !conf.twotime) log_info("--gpus %d: this job (dense frames or --frameout) runs on one GPU", fl.gpus);
    if (sharded) {
        try {
            int T = xpcs_delay_schedule(frames, conf.dpl, nullptr, nullptr, 0);
            XpcsShardPlan plan;
            if (int rc = xpcs_plan_shard(&prm, &plan, nullptr, 0)) {
                fprintf(stderr, "corr: xpcs_plan_shard failed (%d): %s\n", rc, xpcs_last_error(nullptr));
                return 3;
            }
            const int S = plan.n_static, Q = plan.n_dynamic;
            SparseInput in;
            {
                Scope sc("Reading input");
                if (fl.ufxc) load_ufxc(conf, frames, in);
                else if (fl.hdf5) load_hdf5_stack(conf, frames, pixels, in);
                else if (fl.rigaku) load_rigaku(conf, frames, pixels, in);
                else load_imm_sparse(conf, frames, in);
            }
            FilterSums sums;
            std::vector<float> G2, IP, IF, g2((size_t)T * Q), se((size_t)T * Q);
            log_info("Sharding over %d GPUs (frame slabs in, pixel rows out)", fl.gpus);
            if (int rc = run_sharded(fl, prm, in, fl.gpus, sums, T, Q, pixels, fl.g2out ? &G2 : nullptr, fl.g2out ? &IP : nullptr,
                                     fl.g2out ? &IF : nullptr, g2, se))
                return rc;
            const int windows = frames / prm.static_window;
            file.put(out + "/pixelSum", Type::F32, {uy, ux}, sums.pixel_sum.data());
            file.put(out + "/frameSum", Type::F32, {2, (uint64_t)frames}, sums.frame_sum.data());
            file.put(out + "/partition-mean-total", Type::F32, {1, (uint64_t)S}, sums.pm_total.data());
            file.put(out + "/partition-mean-partial", Type::F32, {(uint64_t)windows, (uint64_t)S}, sums.pm_partial.data());
            file.put(out + "/partition_norm_factor", Type::F32, {1, 1}, &conf.norm_factor);
            std::vector<double> ck(2 * (size_t)real_frames, 0.0), tk(2 * (size_t)real_frames, 0.0);
            for (int i = 0; i < real_frames && i < sums.raw_seen; i++) {
                ck[i] = sums.clock[i];
                ck[real_frames + i] = sums.clock[sums.raw_seen + i];
                tk[i] = sums.ticks[i];
                tk[real_frames + i] = sums.ticks[sums.raw_seen + i];
            }
            if (fl.rigaku)
                for (int i = 0; i < real_frames; i++) {
                    ck[i] = tk[i] = (double)i;
                    ck[real_frames + i] = tk[real_frames + i] = 0.0;
                }
            file.put(out + "/timestamp_clock", Type::F64, {2, (uint64_t)real_frames}, ck.data());
            file.put(out + "/timestamp_tick", Type::F64, {2, (uint64_t)real_frames}, tk.data());
            std::vector<int32_t> lv(T), tv(T);
            xpcs_delay_schedule(frames, conf.dpl, lv.data(), tv.data(), T);
            std::vector<float> tau(T);
            for (int i = 0; i < T; i++) tau[i] = (float)tv[i];
            file.put(out + "/tau", Type::F32, {1, (uint64_t)T}, tau.data());
            file.put(out + "/norm-0-g2", Type::F32, {(uint64_t)T, (uint64_t)Q}, g2.data());
            file.put(out + "/norm-0-stderr", Type::F32, {(uint64_t)T, (uint64_t)Q}, se.data());
            if (fl.g2out) {
                Scope sc("Writing G2s, IPs and IFs");
                file.put(out + "/G2", Type::F32, {(uint64_t)T, (uint64_t)pixels}, G2.data());
                file.put(out + "/IP", Type::F32, {(uint64_t)T, (uint64_t)pixels}, IP.data());
                file.put(out + "/IF", Type::F32, {(uint64_t)T, (uint64_t)pixels}, IF.data());
            }
            Scope sc("Writing results");
            file.save(result_file);
        } catch (const std::exception &e) {
            fprintf(stderr, "corr: %s\n", e.what());
            return 1;
        }
        return 0;
    }

    // The handle first, then the input: creating the CUDA context on a second thread while this one faults in a few
    // hundred MB of file and buffers was measured slower than doing one after the other (both sides serialise on the
    // process's memory-map lock: 2.2 s against 1.8 s for the 731 MB file of the 1-Mpixel configuration).
    xpcs_handle h = nullptr;
    int create_rc = xpcs_create(&prm, fl.device, &h);
    std::string create_err = create_rc ? xpcs_last_error(nullptr) : "";
    std::thread creator;
    auto join_creator = [&]() {
        if (create_rc) fprintf(stderr, "corr: xpcs_create failed (%d): %s\n", create_rc, create_err.c_str());
        return create_rc == 0;
    };
    if (!join_creator()) return 3;
    const bool sparse_input = fl.ufxc || fl.hdf5 || fl.rigaku || imm_reader->sparse();
    if (fl.stream_frames > 0 && !sparse_input) {
        fprintf(stderr, "corr: --stream_frames takes sparse input (compressed IMM, --ufxc, --rigaku, --hdf5)\n");
        xpcs_destroy(h);
        return 2;
    }
    SparseInput in;
    std::chrono::steady_clock::time_point t_load = std::chrono::steady_clock::now();
    try {
        if (fl.ufxc) load_ufxc(conf, frames, in);
        else if (fl.hdf5) load_hdf5_stack(conf, frames, pixels, in);
        else if (fl.rigaku) load_rigaku(conf, frames, pixels, in);
        else if (sparse_input && fl.stream_frames <= 0) load_imm_sparse(conf, frames, in);  // (a streamed IMM file is read chunk by chunk below)
    } catch (const std::exception &e) {
        fprintf(stderr, "corr: %s\n", e.what());
        xpcs_destroy(h);
        return 1;
    }
    XpcsInfo info;
    CHECK(xpcs_get_info(h, &info));
    const int T = info.n_delays, S = info.n_static, Q = info.n_dynamic;

    try {
        bool had_dark = false;
        {
            // the stage line covers reading the input too, as the reference's "Loading data" does
            struct LoadScope {
                std::chrono::steady_clock::time_point t0;
                ~LoadScope()
                {
                    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                    if (ms < 1000.0) log_info("Loading data took %.0f ms", ms);
                    else log_info("Loading data took %.3f s", ms / 1e3);
                }
            } sc{t_load};
            if (sparse_input) {
                // one push of the whole frame range: large pushes of plain photon counts are cut into chunks that are
                // ingested while the next chunk crosses PCIe
                if (fl.stream_frames > 0 && (fl.ufxc || fl.hdf5 || fl.rigaku)) {
                    // these readers decode the whole file at once (event words in file order, a frame stack): the
                    // device still holds one chunk at a time, the library cuts the push
                    CHECK(xpcs_stream_begin(h, fl.stream_frames));
                    CHECK(xpcs_stream_push_sparse(h, in.idxp(), in.valp(), in.offs.data(), in.clock.data(), in.ticks.data(), in.raw_frames()));
                } else if (fl.stream_frames > 0) {
                    // a compressed IMM file is walked chunk by chunk: one chunk of frames on the host (page-locked, so
                    // that it crosses PCIe at the link rate), one on the device, whatever the length of the file
                    CHECK(xpcs_stream_begin(h, fl.stream_frames));
                    const int frame_from = conf.frame_start_todo - 1;  // main.cpp:241-245
                    xpcs_host::SparseImmStream src(conf.imm_path, frame_from > 0 ? frame_from : 0);
                    const int K = fl.stream_frames;
                    struct Pinned {
                        void *p = nullptr;
                        size_t n = 0;
                        bool pinned = false;
                        ~Pinned() { drop(); }
                        void drop()
                        {
                            if (p && pinned) xpcs_host_free(p);
                            else free(p);
                            p = nullptr;
                        }
                        void *need(size_t bytes)
                        {
                            if (bytes <= n && p) return p;
                            drop();
                            n = bytes + bytes / 4 + 64;
                            p = xpcs_host_alloc(n);
                            pinned = p != nullptr;
                            if (!p) p = malloc(n);
                            if (!p) throw std::runtime_error("out of memory for a chunk of frames");
                            return p;
                        }
                    } bi, bv;
                    std::vector<int64_t> offs((size_t)K + 1);
                    std::vector<double> ck((size_t)K), tk((size_t)K);
                    for (int done = 0; done < frames;) {
                        const int n = std::min(K, frames - done);
                        const int64_t ne = src.peek_events(n);
                        int32_t *ci = static_cast<int32_t *>(bi.need(sizeof(int32_t) * ((size_t)ne + 8)));
                        int16_t *cv = static_cast<int16_t *>(bv.need(sizeof(int16_t) * ((size_t)ne + 8)));
                        src.next(n, ci, cv, offs.data(), ck.data(), tk.data());
                        CHECK(xpcs_stream_push_sparse(h, ci, cv, offs.data(), ck.data(), tk.data(), n));
                        done += n;
                    }
                } else
                CHECK(xpcs_push_sparse(h, in.idxp(), in.valp(), in.offs.data(), in.clock.data(), in.ticks.data(), in.raw_frames()));
            } else {
                xpcs_host::ImmReader &reader = *imm_reader;
                xpcs_host::ImmBatch b;
                int r = 0;
                if (conf.darks > 0) {  // main.cpp:227-239: darks come from the file start
                    reader.next(conf.darks, b, pixels);
                    CHECK(xpcs_set_dark(h, b.val.data(), conf.darks));
                    had_dark = true;
                    r += conf.darks;
                }
                const int frame_from = conf.frame_start_todo - 1;
                if (frame_from > 0 && r < frame_from) reader.skip(frame_from - r);  // main.cpp:241-245
                const int64_t raw_todo = (int64_t)frames * raw_block(conf);
                const int chunk = std::max(1, (int)((256ll << 20) / ((int64_t)pixels * 2)));
                for (int64_t done = 0; done < raw_todo;) {
                    const int n = (int)std::min<int64_t>(chunk, raw_todo - done);
                    reader.next(n, b, pixels);
                    CHECK(xpcs_push_dense(h, b.val.data(), b.clock.data(), b.ticks.data(), n));
                    done += n;
                }
            }
            std::vector<float> pixel_sum(pixels), frame_sum(2 * (size_t)frames), pm_total(S > 0 ? S : 1);
            const int windows = frames / prm.static_window;
            std::vector<float> pm_partial((size_t)std::max(windows, 1) * std::max(S, 1));
            if (fl.stream_frames > 0 && sparse_input)
                CHECK(xpcs_stream_finish(h, pixel_sum.data(), frame_sum.data(), pm_total.data(), pm_partial.data()));
            else
                CHECK(xpcs_finish_ingest(h, pixel_sum.data(), frame_sum.data(), pm_total.data(), pm_partial.data()));
            CHECK(xpcs_get_info(h, &info));
            const int raw_seen = info.raw_frames_seen;
            std::vector<double> clock(2 * (size_t)raw_seen), ticks(2 * (size_t)raw_seen);
            CHECK(xpcs_get_timestamps(h, clock.data(), ticks.data()));
            // result datasets of main.cpp:345-426 (names, shapes and types of SURVEY.md B.5)
            file.put(out + "/pixelSum", Type::F32, {uy, ux}, pixel_sum.data());
            file.put(out + "/frameSum", Type::F32, {2, (uint64_t)frames}, frame_sum.data());
            file.put(out + "/partition-mean-total", Type::F32, {1, (uint64_t)S}, pm_total.data());
            file.put(out + "/partition-mean-partial", Type::F32, {(uint64_t)windows, (uint64_t)S}, pm_partial.data());
            file.put(out + "/partition_norm_factor", Type::F32, {1, 1}, &conf.norm_factor);
            // the reference sizes these by getRealFrameTodoCount() (main.cpp:399-411)
            std::vector<double> ck(2 * (size_t)real_frames, 0.0), tk(2 * (size_t)real_frames, 0.0);
            for (int i = 0; i < real_frames && i < raw_seen; i++) {
                ck[i] = clock[i];
                ck[real_frames + i] = clock[raw_seen + i];
                tk[i] = ticks[i];
                tk[real_frames + i] = ticks[raw_seen + i];
            }
            if (fl.rigaku) {
                // the Rigaku reader hands main() arrays of `frames` doubles (io/rigaku.cpp:91-92, 184-185: the
                // output frame number), which main() writes as [2][frames] (main.cpp:399-411): the first row is
                // 0 .. frames-1, the second is whatever lies behind the array.  Row 0 as the reference, row 1 zero.
                for (int i = 0; i < real_frames; i++) {
                    ck[i] = tk[i] = (double)i;
                    ck[real_frames + i] = tk[real_frames + i] = 0.0;
                }
            }
            file.put(out + "/timestamp_clock", Type::F64, {2, (uint64_t)real_frames}, ck.data());
            file.put(out + "/timestamp_tick", Type::F64, {2, (uint64_t)real_frames}, tk.data());
        }
        if (fl.darkout && had_dark) {  // main.cpp:459-477
            std::vector<double> avg(pixels), sd(pixels);
            CHECK(xpcs_get_dark(h, avg.data(), sd.data()));
            file.put(out + "/DarkAvg", Type::F64, {uy, ux}, avg.data());
            file.put(out + "/DarkStd", Type::F64, {uy, ux}, sd.data());
        }
        if (fl.frameout > 0 && fl.frameout < frames) {  // main.cpp:276-310
            std::vector<float> fr((size_t)fl.frameout * pixels);
            CHECK(xpcs_get_frames(h, fl.frameout, fr.data()));
            // the reference declares (height, width, N) for a buffer it fills as [N][pixels] (h5_result.cpp:169-226)
            file.put(out + "/frames_out", Type::F32, {uy, ux, (uint64_t)fl.frameout}, fr.data());
        }
        if (conf.twotime) {
            Scope sc("Computing G2 TwoTimes");
            std::vector<int> bins;  // the listed dynamic partitions that exist, ascending (std::map order)
            for (int q = 1; q <= Q; q++)
                for (int want : conf.qphi_bins)
                    if (want == q) {
                        bins.push_back(q);
                        break;
                    }
            const bool average = conf.smoothing_filter == "Average";
            const int F = frames, w = conf.wsize, partials = std::max((F - w) / w, 0);
            const size_t B = bins.size();
            const size_t sg_cols = average ? 1 : (size_t)F;
            std::vector<float> g2full((size_t)F * B), g2part((size_t)w * partials * B);
            std::vector<float> sgall;   // rows: one per dynamic bin (symmetric) or per static bin of the bins (StaticMap)
            // Dynamic partitions are independent (corr.cpp:799 loops over them): with --gpus N every GPU ingests the
            // input and takes every N-th listed partition ("one dynamic bin per GPU").
            struct BinOut {
                int rc = 0, sg_rows = 0;
                std::string err;
                std::vector<float> C, gf, gp, sg;
            };
            std::vector<BinOut> outs(B);
            auto process = [&](xpcs_handle hh, size_t first, size_t stride) {
                for (size_t k = first; k < B; k += stride) {
                    BinOut &o = outs[k];
                    o.C.resize((size_t)F * F);
                    o.gf.resize(F);
                    o.gp.resize((size_t)std::max(w * partials, 1));
                    o.sg.resize((size_t)std::max(S, 1) * sg_cols);
                    o.rc = xpcs_twotime_sg(hh, bins[k], w, method, average ? 1 : 0, o.C.data(), o.gf.data(), o.gp.data(), o.sg.data(),
                                           &o.sg_rows);
                    if (o.rc) {
                        o.err = xpcs_last_error(hh);
                        std::vector<float>().swap(o.C);
                    }
                }
            };
            const int n_tt = (fl.gpus > 1 && sparse_input) ? (int)std::min<size_t>((size_t)fl.gpus, std::max<size_t>(B, 1)) : 1;
            if (fl.gpus > 1 && n_tt > 1) log_info("Two-time: %zu dynamic partitions over %d GPUs", B, n_tt);
            std::vector<std::thread> helpers;
            std::vector<std::string> helper_err((size_t)n_tt);
            for (int r = 1; r < n_tt; r++)
                helpers.emplace_back([&, r]() {
                    xpcs_handle hr = nullptr;
                    int rc = xpcs_create(&prm, fl.device + r, &hr);
                    if (rc) {
                        helper_err[(size_t)r] = std::string("xpcs_create: ") + xpcs_last_error(nullptr);
                        return;
                    }
                    rc = xpcs_push_sparse(hr, in.idxp(), in.valp(), in.offs.data(), in.clock.data(), in.ticks.data(), in.raw_frames());
                    if (!rc) rc = xpcs_finish_ingest(hr, nullptr, nullptr, nullptr, nullptr);
                    if (rc) helper_err[(size_t)r] = std::string("ingest: ") + xpcs_last_error(hr);
                    else process(hr, (size_t)r, (size_t)n_tt);
                    xpcs_destroy(hr);
                });
            process(h, 0, (size_t)n_tt);
            for (auto &t : helpers) t.join();
            for (int r = 1; r < n_tt; r++)
                if (!helper_err[(size_t)r].empty()) {
                    fprintf(stderr, "corr: GPU %d: %s\n", fl.device + r, helper_err[(size_t)r].c_str());
                    return 3;
                }
            size_t b = 0;
            for (size_t k = 0; k < B; k++) {
                BinOut &o = outs[k];
                const int q = bins[k];
                if (o.rc == XPCS_E_ARG) continue;  // partition without pixels: the reference skips it too
                if (o.rc) {
                    fprintf(stderr, "corr: xpcs_twotime failed (%d): %s\n", o.rc, o.err.c_str());
                    return 3;
                }
                char name[64];
                snprintf(name, sizeof(name), "/C2T_all/g2_%05d", q);
                // one chunk covering the matrix, deflate level 6: the reference's storage of these datasets
                // (write2DData(..., compression = true), h5_result.cpp:140-152; corr.cpp:883-890)
                h5lite::Dataset &c2t = file.put(out + name, Type::F32, {(uint64_t)F, (uint64_t)F}, o.C.data());
                std::vector<float>().swap(o.C);
                if (!fl.nocompress && (uint64_t)F * F * 4 < (1ull << 32)) {
                    c2t.chunk = {(uint64_t)F, (uint64_t)F};
                    c2t.deflate_level = 6;
                }
                for (int f = 0; f < F; f++) g2full[(size_t)f * B + b] = o.gf[f];
                for (int d = 0; d < w; d++)
                    for (int p = 0; p < partials; p++) g2part[((size_t)d * partials + p) * B + b] = o.gp[(size_t)d * partials + p];
                sgall.insert(sgall.end(), o.sg.begin(), o.sg.begin() + (size_t)o.sg_rows * sg_cols);
                b++;
            }
            file.put(out + "/sg", Type::F32, {(uint64_t)(sgall.size() / sg_cols), (uint64_t)sg_cols}, sgall.data());
            file.put(out + "/g2full", Type::F32, {(uint64_t)F, (uint64_t)B}, g2full.data());
            file.put(out + "/g2partials", Type::F32, {(uint64_t)w, (uint64_t)partials, (uint64_t)B}, g2part.data());
        } else {
            std::vector<int32_t> lv(T), tv(T);
            xpcs_delay_schedule(frames, conf.dpl, lv.data(), tv.data(), T);
            std::vector<float> tau(T);
            for (int i = 0; i < T; i++) tau[i] = (float)tv[i];
            file.put(out + "/tau", Type::F32, {1, (uint64_t)T}, tau.data());
            std::vector<float> G2, IP, IF;
            {
                Scope sc("Computing G2 MultiTau");
                if (fl.g2out) {
                    G2.resize((size_t)T * pixels);
                    IP.resize((size_t)T * pixels);
                    IF.resize((size_t)T * pixels);
                    CHECK(xpcs_multitau(h, G2.data(), IP.data(), IF.data()));
                } else CHECK(xpcs_multitau(h, nullptr, nullptr, nullptr));
            }
            {
                Scope sc("Normalizing Data");
                std::vector<float> g2((size_t)T * Q), se((size_t)T * Q);
                CHECK(xpcs_normalize(h, g2.data(), se.data()));
                file.put(out + "/norm-0-g2", Type::F32, {(uint64_t)T, (uint64_t)Q}, g2.data());
                file.put(out + "/norm-0-stderr", Type::F32, {(uint64_t)T, (uint64_t)Q}, se.data());
            }
            if (fl.g2out) {
                Scope sc("Writing G2s, IPs and IFs");
                file.put(out + "/G2", Type::F32, {(uint64_t)T, (uint64_t)pixels}, G2.data());
                file.put(out + "/IP", Type::F32, {(uint64_t)T, (uint64_t)pixels}, IP.data());
                file.put(out + "/IF", Type::F32, {(uint64_t)T, (uint64_t)pixels}, IF.data());
            }
        }
        {
            Scope sc("Writing results");
            file.save(result_file);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "corr: %s\n", e.what());
        xpcs_destroy(h);
        return 1;
    }
    xpcs_destroy(h);
    return 0;
}
