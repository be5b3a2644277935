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
#include <string>
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
    std::string imm, inpath, outpath, exchange, entry = "/xpcs", config;
    int device = 0;
    int frameout = 0;
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
            else if (name == "device") f.device = atoi(need().c_str());
            else if (name == "no_compat") f.no_compat = true;
            else if (name == "frame_threading" || name == "noframe_threading") {
            }  // the two-time contraction has one (tensor-core) path
            else if (name == "frameout") f.frameout = atoi(need().c_str());
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
        else if (m == "staticmap") {
            fprintf(stderr, "corr: smoothing_method StaticMap is not built yet (symmetric is)\n");
            return 1;
        } else {
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
    if (fl.rigaku) {
        // the Rigaku reader plays the Filter stage itself and counts the static windows its own way
        // (io/rigaku.cpp:190-193); its stride / average modes are not covered here
        if (conf.stride > 1 || conf.avg > 1) {
            fprintf(stderr, "corr: --rigaku with stride_frames / avg_frames > 1 is outside the scope of this build\n");
            return 2;
        }
        prm.compat_flags |= XPCS_COMPAT_LATE_WINDOW;
    }
    prm.lld = conf.lld;
    prm.sigma = conf.sigma;
    prm.dqmap = conf.dqmap.data();
    prm.sqmap = conf.sqmap.data();
    prm.flatfield = conf.flatfield_enabled ? conf.flatfield.data() : nullptr;
    prm.shard_index = 0;
    prm.shard_count = 1;
    xpcs_handle h = nullptr;
    if (int rc = xpcs_create(&prm, fl.device, &h)) {
        fprintf(stderr, "corr: xpcs_create failed (%d): %s\n", rc, xpcs_last_error(nullptr));
        return 3;
    }
    XpcsInfo info;
    CHECK(xpcs_get_info(h, &info));
    const int T = info.n_delays, S = info.n_static, Q = info.n_dynamic;
    const std::string out = conf.output_path;
    const uint64_t uy = (uint64_t)conf.ydim, ux = (uint64_t)conf.xdim;

    try {
        bool had_dark = false;
        {
            Scope sc("Loading data");
            if (fl.ufxc) {
                // --ufxc (main.cpp:206-207; io/ufxc.cpp:59-153): a stream of 32-bit event words -- frame counter
                // in bits 31..21 (11 bits, unwrapped by +-2048 when it jumps by more than 2000; the first word
                // is frame 0), count in bits 16..15, column-major pixel in bits 14..0.  Frames come out in
                // file order, a missing frame is an empty frame, the reader's SkipFrames does nothing (the
                // frame range always starts at the first frame), clock = ticks = frame number.
                FILE *fp = fopen(conf.imm_path.c_str(), "rb");
                if (!fp) throw std::runtime_error("cannot open " + conf.imm_path);
                std::vector<uint32_t> words;
                {
                    uint32_t buf[4096];
                    size_t got;
                    while ((got = fread(buf, sizeof(uint32_t), 4096, fp)) > 0) words.insert(words.end(), buf, buf + got);
                    fclose(fp);
                }
                int block = conf.stride > 1 ? (int)conf.stride : (int)conf.avg;      // main.cpp:258-261
                if (conf.stride > 1 && conf.avg > 1) block = (int)(conf.stride * conf.avg);
                const int64_t raw_todo = (int64_t)frames * block;
                std::vector<int64_t> count((size_t)raw_todo + 1, 0);
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
                std::vector<int32_t> idx((size_t)count[(size_t)raw_todo] + 1);
                std::vector<int16_t> val((size_t)count[(size_t)raw_todo] + 1);
                {
                    std::vector<int64_t> cur(count.begin(), count.end() - 1);
                    const uint32_t H = (uint32_t)conf.ydim, W = (uint32_t)conf.xdim;
                    for (size_t i = 0; i < words.size(); i++) {
                        if (frame_of[i] < 0) continue;
                        const uint32_t pix = words[i] & 0x7fffu;
                        const int64_t at = cur[(size_t)frame_of[i]]++;  // file order inside a frame
                        idx[(size_t)at] = (int32_t)((pix % H) * W + pix / H);
                        val[(size_t)at] = (int16_t)((words[i] >> 15) & 0x3u);
                    }
                }
                std::vector<double> stamp((size_t)raw_todo);
                for (int64_t f = 0; f < raw_todo; f++) stamp[(size_t)f] = (double)f;
                CHECK(xpcs_push_sparse(h, idx.data(), val.data(), count.data(), stamp.data(), stamp.data(), (int)raw_todo));
            } else if (fl.hdf5) {
                // --hdf5 (main.cpp:211-216; io/hdf5.cpp:62-228): /entry/data/data, uint16 or uint32 [frames][.][.];
                // every non-zero sample of a frame is an event at its linear index, the frames before
                // data_begin_todo are skipped, clock = ticks = frame number in the stack.
                const h5lite::File stack = h5lite::File::load(conf.imm_path);
                const h5lite::Dataset &d = stack.dataset("/entry/data/data");
                if (d.dims.size() != 3 || (d.type != Type::U16 && d.type != Type::U32))
                    throw std::runtime_error("/entry/data/data must be a 3-d uint16 or uint32 dataset");
                const uint64_t nfr = d.dims[0], per = d.dims[1] * d.dims[2];
                if (per != (uint64_t)pixels) throw std::runtime_error("/entry/data/data: frame size differs from the detector");
                int block = conf.stride > 1 ? (int)conf.stride : (int)conf.avg;      // main.cpp:258-261
                if (conf.stride > 1 && conf.avg > 1) block = (int)(conf.stride * conf.avg);
                const int64_t raw_todo = (int64_t)frames * block;
                const int64_t first = conf.frame_start_todo > 1 ? conf.frame_start_todo - 1 : 0;  // main.cpp:241-245
                if ((uint64_t)(first + raw_todo) > nfr) throw std::runtime_error("/entry/data/data holds too few frames");
                std::vector<int32_t> idx;
                std::vector<int16_t> val;
                std::vector<int64_t> offs(1, 0);
                std::vector<double> stamp;
                const uint16_t *p16 = reinterpret_cast<const uint16_t *>(d.data.data());
                const uint32_t *p32 = reinterpret_cast<const uint32_t *>(d.data.data());
                for (int64_t f = first; f < first + raw_todo; f++) {
                    for (uint64_t i = 0; i < per; i++) {
                        const uint32_t v = d.type == Type::U16 ? p16[(uint64_t)f * per + i] : p32[(uint64_t)f * per + i];
                        if (v == 0) continue;
                        if (v > 32767u) throw std::runtime_error("/entry/data/data: a sample above 32767 does not fit the int16 event payload");
                        idx.push_back((int32_t)i);
                        val.push_back((int16_t)v);
                    }
                    offs.push_back((int64_t)idx.size());
                    stamp.push_back((double)f);
                }
                idx.push_back(0);
                val.push_back(0);
                CHECK(xpcs_push_sparse(h, idx.data(), val.data(), offs.data(), stamp.data(), stamp.data(), (int)raw_todo));
            } else if (fl.rigaku) {
                // --rigaku (main.cpp:208-210; io/rigaku.cpp:139-267, stride = average = 1): 64-bit event words,
                // frame number in bits 63..40, column-major pixel in bits 35..16, count in bits 10..0.  Words of
                // frames up to data_begin_todo - 1 are skipped; an output frame is closed when the frame number
                // changes, so frames without events vanish (a run that does not start at data_begin_todo opens
                // with an empty output frame); reading stops when `frames` output frames are closed; masked
                // pixels are dropped after the frame bookkeeping; the frame still open at the end of the file
                // counts only if it holds more than one pixel; clock = ticks = output frame number.
                FILE *fp = fopen(conf.imm_path.c_str(), "rb");
                if (!fp) throw std::runtime_error("cannot open " + conf.imm_path);
                const uint64_t start = (uint64_t)(conf.frame_start_todo > 0 ? conf.frame_start_todo - 1 : 0);
                const uint32_t H = (uint32_t)conf.ydim, W = (uint32_t)conf.xdim;
                std::vector<int32_t> idx;
                std::vector<int16_t> val;
                std::vector<int64_t> offs(1, 0);
                uint64_t open_frame = start + 1;
                size_t open_at = 0;  // first event of the frame still open
                bool stop = false;
                std::vector<uint64_t> buf(1 << 15);
                size_t got;
                while (!stop && (got = fread(buf.data(), sizeof(uint64_t), buf.size(), fp)) > 0) {
                    for (size_t i = 0; i < got; i++) {
                        const uint64_t wd = buf[i];
                        const uint64_t fr = (wd >> 40) & 0xffffffffull;
                        if (fr <= start) continue;
                        if ((int64_t)offs.size() - 1 >= frames) {
                            stop = true;
                            break;
                        }
                        if (fr != open_frame) {
                            offs.push_back((int64_t)idx.size());
                            open_frame = fr;
                            open_at = idx.size();
                        }
                        uint32_t pix = (uint32_t)((wd >> 16) & 0xfffffu);
                        pix = (pix % H) * W + pix / H;
                        if (pix >= (uint32_t)pixels || conf.dqmap[pix] < 1 || conf.sqmap[pix] < 1) continue;
                        idx.push_back((int32_t)pix);
                        val.push_back((int16_t)(wd & 0x7ffu));
                    }
                }
                fclose(fp);
                if ((int64_t)offs.size() - 1 < frames) {  // the open frame: kept if it has more than one distinct pixel
                    bool two = false;
                    for (size_t k = open_at + 1; k < idx.size() && !two; k++) two = idx[k] != idx[open_at];
                    if (two) offs.push_back((int64_t)idx.size());
                    else {
                        idx.resize(open_at);
                        val.resize(open_at);
                    }
                } else {  // `frames` closed: whatever was collected for the next one is dropped
                    idx.resize((size_t)offs.back());
                    val.resize((size_t)offs.back());
                }
                while ((int64_t)offs.size() - 1 < frames) offs.push_back((int64_t)idx.size());
                std::vector<double> stamp((size_t)frames);
                for (int f = 0; f < frames; f++) stamp[(size_t)f] = (double)f;
                idx.push_back(0);
                val.push_back(0);
                CHECK(xpcs_push_sparse(h, idx.data(), val.data(), offs.data(), stamp.data(), stamp.data(), frames));
            } else {
                xpcs_host::ImmReader reader(conf.imm_path);
                xpcs_host::ImmBatch b;
                int r = 0;
                if (!reader.sparse() && conf.darks > 0) {  // main.cpp:227-239: darks come from the file start
                    reader.next(conf.darks, b, pixels);
                    CHECK(xpcs_set_dark(h, b.val.data(), conf.darks));
                    had_dark = true;
                    r += conf.darks;
                }
                const int frame_from = conf.frame_start_todo - 1;
                if (frame_from > 0 && r < frame_from) reader.skip(frame_from - r);  // main.cpp:241-245
                int block = conf.stride > 1 ? (int)conf.stride : (int)conf.avg;      // main.cpp:258-261
                if (conf.stride > 1 && conf.avg > 1) block = (int)(conf.stride * conf.avg);
                const int64_t raw_todo = (int64_t)frames * block;
                const int chunk = reader.sparse() ? 4096 : std::max(1, (int)((256ll << 20) / ((int64_t)pixels * 2)));
                for (int64_t done = 0; done < raw_todo;) {
                    const int n = (int)std::min<int64_t>(chunk, raw_todo - done);
                    reader.next(n, b, pixels);
                    if (reader.sparse())
                        CHECK(xpcs_push_sparse(h, b.idx.data(), b.val.data(), b.offsets.data(), b.clock.data(), b.ticks.data(), n));
                    else CHECK(xpcs_push_dense(h, b.val.data(), b.clock.data(), b.ticks.data(), n));
                    done += n;
                }
            }
            std::vector<float> pixel_sum(pixels), frame_sum(2 * (size_t)frames), pm_total(S > 0 ? S : 1);
            const int windows = frames / prm.static_window;
            std::vector<float> pm_partial((size_t)std::max(windows, 1) * std::max(S, 1));
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
            std::vector<float> C((size_t)F * F), gf(F), gp((size_t)std::max(w * partials, 1)), sg(average ? 1 : F);
            std::vector<float> g2full((size_t)F * B), g2part((size_t)w * partials * B), sgall((average ? 1 : (size_t)F) * B);
            size_t b = 0;
            for (int q : bins) {
                int rc = xpcs_twotime(h, q, w, method, average ? 1 : 0, C.data(), gf.data(), gp.data(), sg.data());
                if (rc == XPCS_E_ARG) continue;  // partition without pixels: the reference skips it too
                if (rc) {
                    fprintf(stderr, "corr: xpcs_twotime failed (%d): %s\n", rc, xpcs_last_error(h));
                    return 3;
                }
                char name[64];
                snprintf(name, sizeof(name), "/C2T_all/g2_%05d", q);
                file.put(out + name, Type::F32, {(uint64_t)F, (uint64_t)F}, C.data());
                for (int f = 0; f < F; f++) g2full[(size_t)f * B + b] = gf[f];
                for (int d = 0; d < w; d++)
                    for (int p = 0; p < partials; p++) g2part[((size_t)d * partials + p) * B + b] = gp[(size_t)d * partials + p];
                for (size_t i = 0; i < sg.size(); i++) sgall[b * sg.size() + i] = sg[i];
                b++;
            }
            file.put(out + "/sg", Type::F32, {(uint64_t)B, (uint64_t)sg.size()}, sgall.data());
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
            file.save(fl.config);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "corr: %s\n", e.what());
        xpcs_destroy(h);
        return 1;
    }
    xpcs_destroy(h);
    return 0;
}
