// h5lite.cpp -- see h5lite.h.  Format notes are in SURVEY.md Appendix D.
#include "h5lite.h"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace h5lite {

static const uint64_t UNDEF = ~0ull;

size_t type_size(Type t)
{
    switch (t) {
    case Type::I8: case Type::U8: return 1;
    case Type::I16: case Type::U16: return 2;
    case Type::I32: case Type::U32: case Type::F32: return 4;
    case Type::I64: case Type::U64: case Type::F64: return 8;
    default: return 1;
    }
}

uint64_t Dataset::count() const
{
    uint64_t n = 1;
    for (uint64_t d : dims) n *= d;
    return n;
}

template <typename T>
static std::vector<T> convert(const Dataset &d)
{
    const uint64_t n = d.count();
    std::vector<T> out(n);
    const uint8_t *p = d.data.data();
    if (d.data.size() < n * d.elem_size) throw Error("dataset shorter than its dataspace");
    for (uint64_t i = 0; i < n; i++) {
        switch (d.type) {
        case Type::I8: out[i] = (T)((const int8_t *)p)[i]; break;
        case Type::U8: out[i] = (T)((const uint8_t *)p)[i]; break;
        case Type::I16: { int16_t v; memcpy(&v, p + 2 * i, 2); out[i] = (T)v; break; }
        case Type::U16: { uint16_t v; memcpy(&v, p + 2 * i, 2); out[i] = (T)v; break; }
        case Type::I32: { int32_t v; memcpy(&v, p + 4 * i, 4); out[i] = (T)v; break; }
        case Type::U32: { uint32_t v; memcpy(&v, p + 4 * i, 4); out[i] = (T)v; break; }
        case Type::I64: { int64_t v; memcpy(&v, p + 8 * i, 8); out[i] = (T)v; break; }
        case Type::U64: { uint64_t v; memcpy(&v, p + 8 * i, 8); out[i] = (T)v; break; }
        case Type::F32: { float v; memcpy(&v, p + 4 * i, 4); out[i] = (T)v; break; }
        case Type::F64: { double v; memcpy(&v, p + 8 * i, 8); out[i] = (T)v; break; }
        default: throw Error("string dataset read as numbers");
        }
    }
    return out;
}

std::vector<int32_t> Dataset::as_i32() const { return convert<int32_t>(*this); }
std::vector<int64_t> Dataset::as_i64() const { return convert<int64_t>(*this); }
std::vector<float> Dataset::as_f32() const { return convert<float>(*this); }
std::vector<double> Dataset::as_f64() const { return convert<double>(*this); }
double Dataset::scalar() const
{
    if (count() < 1) throw Error("empty dataset read as scalar");
    Dataset one = *this;
    one.dims.clear();
    return convert<double>(one)[0];
}
std::string Dataset::as_string() const
{
    if (type != Type::STR) throw Error("numeric dataset read as string");
    size_t n = std::min(elem_size, data.size());
    std::string s((const char *)data.data(), n);
    size_t z = s.find('\0');
    if (z != std::string::npos) s.resize(z);
    while (!s.empty() && s.back() == ' ') s.pop_back();
    return s;
}

// ------------------------------------------------------------------------------------------
// tree access
// ------------------------------------------------------------------------------------------
File::File() { root.is_group = true; }

static std::vector<std::string> split(const std::string &path)
{
    std::vector<std::string> out;
    std::string cur;
    for (char c : path) {
        if (c == '/') {
            if (!cur.empty()) out.push_back(cur);
            cur.clear();
        } else cur.push_back(c);
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}

const Node *File::find(const std::string &path) const
{
    const Node *n = &root;
    for (const std::string &part : split(path)) {
        if (!n->is_group) return nullptr;
        auto it = n->children.find(part);
        if (it == n->children.end()) return nullptr;
        n = it->second.get();
    }
    return n;
}
Node *File::find(const std::string &path) { return const_cast<Node *>(static_cast<const File *>(this)->find(path)); }

const Dataset &File::dataset(const std::string &path) const
{
    const Node *n = find(path);
    if (!n) throw Error("no such dataset: " + path);
    if (n->is_group) throw Error("is a group, not a dataset: " + path);
    return n->ds;
}

Node &File::make_group(const std::string &path)
{
    Node *n = &root;
    for (const std::string &part : split(path)) {
        auto it = n->children.find(part);
        if (it == n->children.end()) {
            std::unique_ptr<Node> g(new Node());
            g->is_group = true;
            it = n->children.emplace(part, std::move(g)).first;
        } else if (!it->second->is_group) throw Error("path component is a dataset: " + part);
        n = it->second.get();
    }
    return *n;
}

Dataset &File::put(const std::string &path, Type t, const std::vector<uint64_t> &dims, const void *data, size_t str_len)
{
    std::vector<std::string> parts = split(path);
    if (parts.empty()) throw Error("empty dataset path");
    std::string leaf = parts.back();
    std::string parent;
    for (size_t i = 0; i + 1 < parts.size(); i++) parent += "/" + parts[i];
    Node &g = make_group(parent);
    std::unique_ptr<Node> n(new Node());
    n->is_group = false;
    n->ds.type = t;
    n->ds.elem_size = t == Type::STR ? str_len : type_size(t);
    n->ds.dims = dims;
    const size_t bytes = (size_t)n->ds.count() * n->ds.elem_size;
    n->ds.data.resize(bytes);
    if (bytes && data) memcpy(n->ds.data.data(), data, bytes);
    Dataset &ref = n->ds;
    g.children[leaf] = std::move(n);  // overwrite in place if it exists (h5_result.cpp:76-103)
    return ref;
}

Dataset &File::put_string(const std::string &path, const std::string &value)
{
    return put(path, Type::STR, {1}, value.c_str(), value.size() + 1);
}

std::vector<std::string> File::list(const std::string &group) const
{
    std::vector<std::string> out;
    const Node *n = find(group);
    if (n && n->is_group)
        for (auto &kv : n->children) out.push_back(kv.first);
    return out;
}

// ------------------------------------------------------------------------------------------
// reader
// ------------------------------------------------------------------------------------------
namespace {

struct Rd {
    std::vector<uint8_t> buf;
    uint64_t base = 0;
    int so = 8, sl = 8;
    int depth = 0;
    std::vector<std::string> *lossy = nullptr, *notes = nullptr;
    std::string path;  // of the node being read

    const uint8_t *at(uint64_t off, uint64_t n) const
    {
        if (off > buf.size() || n > buf.size() - off) throw Error("file truncated or corrupt (read beyond end)");
        return buf.data() + off;
    }
    uint64_t uN(uint64_t off, int n) const
    {
        const uint8_t *p = at(off, n);
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; i--) v = (v << 8) | p[i];
        if (n < 8 && v == ((1ull << (8 * n)) - 1)) return UNDEF;  // undefined address in a narrow field
        return v;
    }
    uint32_t u8(uint64_t o) const { return *at(o, 1); }
    uint32_t u16(uint64_t o) const { const uint8_t *p = at(o, 2); return p[0] | (p[1] << 8); }
    uint32_t u32(uint64_t o) const { const uint8_t *p = at(o, 4); return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
    uint64_t u64(uint64_t o) const { return uN(o, 8); }
    uint64_t off_(uint64_t o) const { return uN(o, so); }
    uint64_t len_(uint64_t o) const { return uN(o, sl); }
    uint64_t abs(uint64_t addr) const { return addr + base; }
};

struct Msg {
    int type;
    uint64_t pos;  // absolute file offset of the message data
    int size;
    int flags;
};

std::vector<Msg> read_object_header(const Rd &r, uint64_t addr)
{
    const uint64_t a = r.abs(addr);
    if (memcmp(r.at(a, 4), "OHDR", 4) == 0) throw Error("version-2 object headers (libver='latest') are not supported");
    if (r.u8(a) != 1) throw Error("unsupported object header version");
    const int nmsgs = r.u16(a + 2);
    const uint64_t hsize = r.u32(a + 8);
    std::vector<std::pair<uint64_t, uint64_t>> blocks;
    blocks.emplace_back(a + 16, hsize);
    std::vector<Msg> out;
    for (size_t b = 0; b < blocks.size() && (int)out.size() < nmsgs; b++) {
        uint64_t p = blocks[b].first, end = blocks[b].first + blocks[b].second;
        while (p + 8 <= end && (int)out.size() < nmsgs) {
            Msg m;
            m.type = r.u16(p);
            m.size = r.u16(p + 2);
            m.flags = r.u8(p + 4);
            m.pos = p + 8;
            r.at(m.pos, m.size);
            if (m.type == 0x0010) blocks.emplace_back(r.abs(r.off_(m.pos)), r.len_(m.pos + r.so));
            out.push_back(m);
            p += 8 + (uint64_t)m.size;
        }
    }
    return out;
}

struct TypeInfo {
    Type type = Type::U8;
    size_t size = 1;
    bool vlen_string = false;
};

TypeInfo parse_datatype(const Rd &r, const Msg &m)
{
    TypeInfo t;
    const int cls = r.u8(m.pos) & 0x0f;
    const int bits0 = r.u8(m.pos + 1);
    t.size = r.u32(m.pos + 4);
    if (cls == 0) {
        if ((bits0 & 1) && t.size > 1) throw Error("big-endian integers are not supported");
        const bool sg = bits0 & 0x08;
        switch (t.size) {
        case 1: t.type = sg ? Type::I8 : Type::U8; break;
        case 2: t.type = sg ? Type::I16 : Type::U16; break;
        case 4: t.type = sg ? Type::I32 : Type::U32; break;
        case 8: t.type = sg ? Type::I64 : Type::U64; break;
        default: throw Error("unsupported integer size");
        }
    } else if (cls == 1) {
        if (bits0 & 1) throw Error("big-endian floats are not supported");
        if (t.size == 4) t.type = Type::F32;
        else if (t.size == 8) t.type = Type::F64;
        else throw Error("unsupported float size");
    } else if (cls == 3) {
        t.type = Type::STR;
    } else if (cls == 9) {
        if ((bits0 & 0x0f) != 1) throw Error("variable-length sequences are not supported (only strings)");
        t.type = Type::STR;
        t.vlen_string = true;
    } else throw Error("unsupported datatype class " + std::to_string(cls));
    return t;
}

std::vector<uint8_t> inflate_chunk(const uint8_t *src, size_t n, size_t expect)
{
    std::vector<uint8_t> out(expect ? expect : 1);
    for (;;) {
        uLongf dl = (uLongf)out.size();
        int rc = uncompress(out.data(), &dl, src, (uLong)n);
        if (rc == Z_OK) {
            out.resize(dl);
            return out;
        }
        if (rc != Z_BUF_ERROR) throw Error("deflate stream corrupt");
        out.resize(out.size() * 2);
    }
}

std::vector<uint8_t> unshuffle(const std::vector<uint8_t> &in, size_t es)
{
    if (es <= 1) return in;
    const size_t n = in.size() / es;
    std::vector<uint8_t> out(in.size());
    for (size_t b = 0; b < es; b++)
        for (size_t i = 0; i < n; i++) out[i * es + b] = in[b * n + i];
    for (size_t i = n * es; i < in.size(); i++) out[i] = in[i];
    return out;
}

struct Filter { int id; };

void read_chunk_tree(const Rd &r, uint64_t addr, int rank, const std::vector<uint64_t> &dims,
                     const std::vector<uint32_t> &cdims, size_t es, const std::vector<Filter> &filters,
                     std::vector<uint8_t> &out)
{
    if (addr == UNDEF) return;  // never written: zeros
    const uint64_t a = r.abs(addr);
    if (memcmp(r.at(a, 4), "TREE", 4) != 0 || r.u8(a + 4) != 1) throw Error("bad chunk B-tree node");
    const int level = r.u8(a + 5);
    const int n = r.u16(a + 6);
    const uint64_t keysz = 8 + 8ull * (rank + 1);
    uint64_t p = a + 8 + 2ull * r.so;
    for (int e = 0; e < n; e++) {
        const uint32_t csize = r.u32(p);
        const uint32_t fmask = r.u32(p + 4);
        std::vector<uint64_t> off(rank);
        for (int d = 0; d < rank; d++) off[d] = r.u64(p + 8 + 8ull * d);
        const uint64_t child = r.off_(p + keysz);
        p += keysz + r.so;
        if (level > 0) {
            read_chunk_tree(r, child, rank, dims, cdims, es, filters, out);
            continue;
        }
        size_t cn = es;
        for (int d = 0; d < rank; d++) cn *= cdims[d];
        std::vector<uint8_t> chunk(r.at(r.abs(child), csize), r.at(r.abs(child), csize) + csize);
        for (int f = (int)filters.size() - 1; f >= 0; f--) {
            if (fmask & (1u << f)) continue;
            if (filters[f].id == 1) chunk = inflate_chunk(chunk.data(), chunk.size(), cn);
            else if (filters[f].id == 2) chunk = unshuffle(chunk, es);
            else throw Error("unsupported chunk filter id " + std::to_string(filters[f].id));
        }
        if (chunk.size() < cn) throw Error("chunk shorter than its extent");
        // copy the chunk into the dataset, clipping at the edges; row-major, last dim contiguous
        std::vector<uint64_t> idx(rank, 0);
        const uint64_t inner = rank ? cdims[rank - 1] : 1;
        const uint64_t rows = cn / es / (inner ? inner : 1);
        for (uint64_t row = 0; row < rows; row++) {
            uint64_t rem = row;
            bool inside = true;
            uint64_t dst = 0;
            for (int d = rank - 2; d >= 0; d--) {
                idx[d] = rem % cdims[d];
                rem /= cdims[d];
            }
            uint64_t stride = 1;
            for (int d = rank - 1; d >= 0; d--) {
                const uint64_t g = off[d] + (d == rank - 1 ? 0 : idx[d]);
                if (d != rank - 1 && g >= dims[d]) inside = false;
                dst += g * stride;
                stride *= dims[d];
            }
            if (!inside) continue;
            uint64_t ncopy = inner;
            if (rank && off[rank - 1] + ncopy > dims[rank - 1]) ncopy = dims[rank - 1] > off[rank - 1] ? dims[rank - 1] - off[rank - 1] : 0;
            if (ncopy) memcpy(out.data() + dst * es, chunk.data() + row * inner * es, ncopy * es);
        }
    }
}

std::string read_vlen_string(const Rd &r, const uint8_t *elem)
{
    uint32_t len;
    memcpy(&len, elem, 4);
    uint64_t gaddr = 0;
    for (int i = r.so - 1; i >= 0; i--) gaddr = (gaddr << 8) | elem[4 + i];
    uint32_t index;
    memcpy(&index, elem + 4 + r.so, 4);
    const uint64_t a = r.abs(gaddr);
    if (memcmp(r.at(a, 4), "GCOL", 4) != 0) throw Error("bad global heap collection");
    const uint64_t csize = r.len_(a + 8);
    uint64_t p = a + 8 + r.sl;
    while (p + 16 <= a + csize) {
        const uint32_t idx = r.u16(p);
        const uint64_t osz = r.len_(p + 8);
        if (idx == 0) break;
        if (idx == index) return std::string((const char *)r.at(p + 8 + r.sl, osz), std::min<uint64_t>(osz, len));
        p += 8 + r.sl + ((osz + 7) & ~7ull);
    }
    throw Error("global heap object not found");
}

void read_node(Rd &r, uint64_t ohdr_addr, Node &node);

// Can an attribute message be copied byte for byte into another file?  Not if the message is shared, or if its
// datatype / dataspace are shared, or if its values live in the global heap (variable-length, reference).
bool attribute_is_self_contained(const Rd &r, const Msg &m, std::string &name)
{
    if (m.size < 8) return false;
    const int ver = r.u8(m.pos);
    const int name_sz = r.u16(m.pos + 2);
    uint64_t p = m.pos + (ver == 3 ? 9 : 8);
    if (ver < 1 || ver > 3 || p + name_sz > m.pos + m.size) return false;
    const char *nm = (const char *)r.at(p, name_sz);
    name.assign(nm, strnlen(nm, name_sz));
    if (m.flags & 0x02) return false;                      // shared message
    if (ver >= 2 && (r.u8(m.pos + 1) & 0x03)) return false;  // shared datatype / dataspace
    p += ver == 1 ? ((uint64_t)name_sz + 7) & ~7ull : (uint64_t)name_sz;
    if (p + 4 > m.pos + m.size) return false;
    const int cls = r.u8(p) & 0x0f;
    return cls == 0 || cls == 1 || cls == 3 || cls == 4 || cls == 5 || cls == 8;
}

void collect_extra(Rd &r, const std::vector<Msg> &msgs, Node &node)
{
    for (const Msg &m : msgs) {
        if (m.type == 0x000C) {
            std::string name;
            if (attribute_is_self_contained(r, m, name)) {
                RawMessage x;
                x.type = (uint16_t)m.type;
                x.flags = (uint8_t)(m.flags & ~0x02);
                x.body.assign(r.at(m.pos, m.size), r.at(m.pos, m.size) + m.size);
                x.body.resize((x.body.size() + 7) & ~7ull, 0);
                node.extra.push_back(std::move(x));
            } else if (r.lossy)
                r.lossy->push_back("attribute '" + name + "' of " + (r.path.empty() ? "/" : r.path) +
                                   " (variable-length, reference or shared: cannot be carried over)");
        } else if (m.type == 0x000D && !(m.flags & 0x02)) {
            RawMessage x;
            x.type = (uint16_t)m.type;
            x.flags = (uint8_t)m.flags;
            x.body.assign(r.at(m.pos, m.size), r.at(m.pos, m.size) + m.size);
            x.body.resize((x.body.size() + 7) & ~7ull, 0);
            node.extra.push_back(std::move(x));
        } else if (m.type == 0x0007 && r.lossy) r.lossy->push_back("external storage of " + r.path);
    }
}

void read_group_entries(Rd &r, uint64_t btree, uint64_t heap, Node &node)
{
    const uint64_t h = r.abs(heap);
    if (memcmp(r.at(h, 4), "HEAP", 4) != 0) throw Error("bad local heap");
    const uint64_t hdata = r.abs(r.off_(h + 8 + 2ull * r.sl));
    const uint64_t t = r.abs(btree);
    if (memcmp(r.at(t, 4), "TREE", 4) != 0 || r.u8(t + 4) != 0) throw Error("bad group B-tree node");
    const int level = r.u8(t + 5);
    const int n = r.u16(t + 6);
    uint64_t p = t + 8 + 2ull * r.so + r.sl;  // skip key 0
    for (int e = 0; e < n; e++) {
        const uint64_t child = r.off_(p);
        p += r.so + r.sl;
        if (level > 0) {
            read_group_entries(r, child, heap, node);
            continue;
        }
        const uint64_t s = r.abs(child);
        if (memcmp(r.at(s, 4), "SNOD", 4) != 0) throw Error("bad symbol node");
        const int ns = r.u16(s + 6);
        const uint64_t esz = 2ull * r.so + 8 + 16;
        for (int i = 0; i < ns; i++) {
            const uint64_t ep = s + 8 + esz * i;
            const uint64_t noff = r.off_(ep);
            const uint64_t oaddr = r.off_(ep + r.so);
            const char *nm = (const char *)r.at(hdata + noff, 1);
            std::string name(nm, strnlen(nm, r.buf.size() - (hdata + noff)));
            std::unique_ptr<Node> c(new Node());
            const std::string outer = r.path;
            r.path = outer + "/" + name;
            read_node(r, oaddr, *c);
            r.path = outer;
            node.children[name] = std::move(c);
        }
    }
}

void read_node(Rd &r, uint64_t ohdr_addr, Node &node)
{
    if (++r.depth > 64) throw Error("group nesting too deep (cycle?)");
    std::vector<Msg> msgs = read_object_header(r, ohdr_addr);
    const Msg *space = nullptr, *dtype = nullptr, *layout = nullptr, *stab = nullptr, *pipeline = nullptr;
    for (const Msg &m : msgs) {
        if (m.type == 0x0001) space = &m;
        else if (m.type == 0x0003) dtype = &m;
        else if (m.type == 0x0008) layout = &m;
        else if (m.type == 0x0011) stab = &m;
        else if (m.type == 0x000B) pipeline = &m;
        else if (m.type == 0x0002) throw Error("new-style (link info) groups are not supported");
    }
    collect_extra(r, msgs, node);
    if (stab) {
        node.is_group = true;
        read_group_entries(r, r.off_(stab->pos), r.off_(stab->pos + r.so), node);
        r.depth--;
        return;
    }
    if (!space || !dtype || !layout) throw Error("object is neither an old-style group nor a simple dataset");
    node.is_group = false;
    Dataset &d = node.ds;
    // dataspace
    const int sver = r.u8(space->pos);
    const int rank = r.u8(space->pos + 1);
    const uint64_t dp = space->pos + (sver == 1 ? 8 : 4);
    d.dims.clear();
    for (int i = 0; i < rank; i++) d.dims.push_back(r.len_(dp + (uint64_t)r.sl * i));
    TypeInfo ti = parse_datatype(r, *dtype);
    d.type = ti.type;
    const size_t disk_es = ti.vlen_string ? (size_t)(8 + r.so) : ti.size;
    d.elem_size = disk_es;
    const uint64_t n = d.count();
    std::vector<uint8_t> raw((size_t)(n * disk_es));
    // layout
    const int lver = r.u8(layout->pos);
    if (lver == 3) {
        const int cls = r.u8(layout->pos + 1);
        if (cls == 0) {
            const int sz = r.u16(layout->pos + 2);
            memcpy(raw.data(), r.at(layout->pos + 4, sz), std::min<size_t>(sz, raw.size()));
        } else if (cls == 1) {
            const uint64_t addr = r.off_(layout->pos + 2);
            if (addr != UNDEF && !raw.empty()) memcpy(raw.data(), r.at(r.abs(addr), raw.size()), raw.size());
        } else if (cls == 2) {
            const int dim = r.u8(layout->pos + 2);
            if (dim != rank + 1) throw Error("chunk dimensionality mismatch");
            const uint64_t bt = r.off_(layout->pos + 3);
            std::vector<uint32_t> cd(dim);
            for (int i = 0; i < dim; i++) cd[i] = r.u32(layout->pos + 3 + r.so + 4ull * i);
            std::vector<Filter> filters;
            if (pipeline) {
                const int pv = r.u8(pipeline->pos);
                const int nf = r.u8(pipeline->pos + 1);
                uint64_t p = pipeline->pos + (pv == 1 ? 8 : 2);
                for (int f = 0; f < nf; f++) {
                    const int id = r.u16(p);
                    int nlen = 0;
                    uint64_t q = p + 2;
                    if (pv == 1 || id >= 256) { nlen = r.u16(q); q += 2; }
                    q += 2;  // flags
                    const int ncd = r.u16(q);
                    q += 2;
                    if (pv == 1) nlen = (nlen + 7) & ~7;
                    q += nlen + 4ull * ncd;
                    if (pv == 1 && (ncd & 1)) q += 4;
                    filters.push_back(Filter{id});
                    p = q;
                }
            }
            read_chunk_tree(r, bt, rank, d.dims, cd, disk_es, filters, raw);
            if (r.notes) r.notes->push_back(r.path + ": chunked" + (filters.empty() ? "" : " + filtered") + " -> contiguous");
        } else throw Error("unsupported layout class");
    } else if (lver == 1 || lver == 2) {
        const int dim = r.u8(layout->pos + 1);
        const int cls = r.u8(layout->pos + 2);
        uint64_t p = layout->pos + 8;
        if (cls == 1) {
            const uint64_t addr = r.off_(p);
            if (addr != UNDEF && !raw.empty()) memcpy(raw.data(), r.at(r.abs(addr), raw.size()), raw.size());
        } else if (cls == 0) {
            p += 4ull * dim;
            const uint32_t sz = r.u32(p);
            memcpy(raw.data(), r.at(p + 4, sz), std::min<size_t>(sz, raw.size()));
        } else if (cls == 2) {
            const uint64_t bt = r.off_(p);
            std::vector<uint32_t> cd(dim);
            for (int i = 0; i < dim; i++) cd[i] = r.u32(p + r.so + 4ull * i);
            if (pipeline) throw Error("filtered chunks with a version-1/2 layout are not supported");
            read_chunk_tree(r, bt, rank, d.dims, cd, disk_es, {}, raw);
        } else throw Error("unsupported layout class");
    } else throw Error("unsupported layout message version");
    if (ti.vlen_string) {
        if (r.notes) r.notes->push_back(r.path + ": variable-length string -> fixed-length string");
        std::string s = n ? read_vlen_string(r, raw.data()) : std::string();
        d.elem_size = s.size() + 1;
        d.dims = {1};
        d.data.assign(s.begin(), s.end());
        d.data.push_back(0);
    } else d.data.swap(raw);
    r.depth--;
}

}  // namespace

File File::load(const std::string &path)
{
    Rd r;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw Error("cannot open " + path);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    r.buf.resize(sz > 0 ? (size_t)sz : 0);
    if (sz > 0 && fread(r.buf.data(), 1, (size_t)sz, f) != (size_t)sz) {
        fclose(f);
        throw Error("short read on " + path);
    }
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    uint64_t sb = UNDEF;
    for (uint64_t off = 0; off + 8 <= r.buf.size(); off = off ? off * 2 : 512)
        if (memcmp(r.buf.data() + off, sig, 8) == 0) {
            sb = off;
            break;
        }
    if (sb == UNDEF) throw Error("not an HDF5 file (no superblock signature): " + path);
    const int ver = r.u8(sb + 8);
    if (ver > 1) throw Error("superblock version " + std::to_string(ver) + " (libver='latest') is not supported");
    r.so = r.u8(sb + 13);
    r.sl = r.u8(sb + 14);
    if ((r.so != 4 && r.so != 8) || (r.sl != 4 && r.sl != 8)) throw Error("unsupported offset/length size");
    uint64_t p = sb + 24 + (ver == 1 ? 4 : 0);
    const uint64_t base_addr = r.uN(p, r.so);
    r.base = base_addr == UNDEF ? sb : base_addr;
    p += 4ull * r.so;
    // root symbol table entry
    const uint64_t root_ohdr = r.off_(p + r.so);
    File out;
    r.lossy = &out.lossy;
    r.notes = &out.notes;
    read_node(r, root_ohdr, out.root);
    if (!out.root.is_group) throw Error("root object is not a group");
    return out;
}

// ------------------------------------------------------------------------------------------
// writer
// ------------------------------------------------------------------------------------------
namespace {

const int kLeafK = 32;   // a symbol node holds up to 2*kLeafK entries
const int kIntK = 16;    // a B-tree node holds up to 2*kIntK children

struct Wr {
    std::vector<uint8_t> b;
    uint64_t alloc(uint64_t n)
    {
        uint64_t at = (b.size() + 7) & ~7ull;
        b.resize(at + n, 0);
        return at;
    }
    void w8(uint64_t o, uint32_t v) { b[o] = (uint8_t)v; }
    void w16(uint64_t o, uint32_t v) { b[o] = v & 0xff; b[o + 1] = (v >> 8) & 0xff; }
    void w32(uint64_t o, uint32_t v) { for (int i = 0; i < 4; i++) b[o + i] = (v >> (8 * i)) & 0xff; }
    void w64(uint64_t o, uint64_t v) { for (int i = 0; i < 8; i++) b[o + i] = (v >> (8 * i)) & 0xff; }
    void bytes(uint64_t o, const void *p, size_t n) { if (n) memcpy(b.data() + o, p, n); }
};

// returns the number of bytes of the datatype message body
size_t datatype_body(const Dataset &d, uint8_t *out)
{
    memset(out, 0, 24);
    const uint32_t size = (uint32_t)(d.type == Type::STR ? d.elem_size : type_size(d.type));
    auto put32 = [&](int o, uint32_t v) { for (int i = 0; i < 4; i++) out[o + i] = (v >> (8 * i)) & 0xff; };
    auto put16 = [&](int o, uint32_t v) { out[o] = v & 0xff; out[o + 1] = (v >> 8) & 0xff; };
    put32(4, size);
    switch (d.type) {
    case Type::F32: case Type::F64: {
        out[0] = 0x11;
        out[1] = 0x20;
        out[2] = (uint8_t)(size * 8 - 1);
        put16(8, 0);
        put16(10, size * 8);
        out[12] = size == 4 ? 23 : 52;
        out[13] = size == 4 ? 8 : 11;
        out[14] = 0;
        out[15] = size == 4 ? 23 : 52;
        put32(16, size == 4 ? 127 : 1023);
        return 20;
    }
    case Type::STR:
        out[0] = 0x13;
        out[1] = 0x01;  // null-padded, ASCII
        return 8;
    default: {
        const bool sg = d.type == Type::I8 || d.type == Type::I16 || d.type == Type::I32 || d.type == Type::I64;
        out[0] = 0x10;
        out[1] = sg ? 0x08 : 0x00;
        put16(8, 0);
        put16(10, size * 8);
        return 12;
    }
    }
}

size_t extra_bytes(const Node &n)
{
    size_t b = 0;
    for (const RawMessage &x : n.extra) b += 8 + x.body.size();
    return b;
}

void write_extra(Wr &w, uint64_t &p, const Node &n)
{
    for (const RawMessage &x : n.extra) {
        w.w16(p, x.type);
        w.w16(p + 2, (uint32_t)x.body.size());
        w.w8(p + 4, x.flags);
        w.bytes(p + 8, x.body.data(), x.body.size());
        p += 8 + x.body.size();
    }
}

std::vector<uint8_t> shuffle_bytes(const std::vector<uint8_t> &in, size_t es)
{
    if (es <= 1) return in;
    const size_t n = in.size() / es;
    std::vector<uint8_t> out(in.size());
    for (size_t i = 0; i < n; i++)
        for (size_t b = 0; b < es; b++) out[b * n + i] = in[i * es + b];
    for (size_t i = n * es; i < in.size(); i++) out[i] = in[i];
    return out;
}

// Chunked storage: every chunk (edge chunks padded with zeros) filtered and written, then a version-1 B-tree of
// node type 1 over them -- one leaf level of up to 2K entries, one internal level above it when needed (K = 32,
// libhdf5's default "indexed storage" rank, which the version-0 superblock implies).  Returns the B-tree address.
uint64_t write_chunks(Wr &w, const Dataset &d, size_t es)
{
    const int rank = (int)d.dims.size();
    const int K = 32;
    const uint64_t keysz = 8 + 8ull * (rank + 1);
    const uint64_t node_bytes = 24 + (2 * K + 1) * keysz + 2 * K * 8ull;
    std::vector<uint64_t> nchunks(rank), cidx(rank, 0);
    uint64_t total = 1, celems = 1;
    for (int i = 0; i < rank; i++) {
        nchunks[i] = (d.dims[i] + d.chunk[i] - 1) / d.chunk[i];
        total *= nchunks[i];
        celems *= d.chunk[i];
    }
    struct Entry { uint64_t addr; uint32_t size; std::vector<uint64_t> off; };
    std::vector<Entry> entries;
    std::vector<uint8_t> buf(celems * es);
    for (uint64_t c = 0; c < total; c++) {
        // gather the chunk (row-major inside the chunk), zero padding beyond the dataset's edge
        std::fill(buf.begin(), buf.end(), 0);
        const uint64_t inner = d.chunk[rank - 1];
        const uint64_t rows = celems / inner;
        for (uint64_t row = 0; row < rows; row++) {
            uint64_t rem = row, src = 0, stride = 1;
            bool inside = true;
            std::vector<uint64_t> g(rank, 0);
            for (int k = rank - 2; k >= 0; k--) {
                g[k] = cidx[k] * d.chunk[k] + rem % d.chunk[k];
                rem /= d.chunk[k];
                if (g[k] >= d.dims[k]) inside = false;
            }
            if (!inside) continue;
            g[rank - 1] = cidx[rank - 1] * d.chunk[rank - 1];
            for (int k = rank - 1; k >= 0; k--) {
                src += g[k] * stride;
                stride *= d.dims[k];
            }
            uint64_t ncopy = inner;
            if (g[rank - 1] + ncopy > d.dims[rank - 1]) ncopy = d.dims[rank - 1] - g[rank - 1];
            memcpy(buf.data() + row * inner * es, d.data.data() + src * es, ncopy * es);
        }
        std::vector<uint8_t> out = d.shuffle ? shuffle_bytes(buf, es) : buf;
        if (d.deflate_level > 0) {
            uLongf bound = compressBound((uLong)out.size());
            std::vector<uint8_t> z(bound);
            if (compress2(z.data(), &bound, out.data(), (uLong)out.size(), d.deflate_level) != Z_OK) throw Error("deflate failed");
            z.resize(bound);
            out.swap(z);
        }
        Entry e;
        e.size = (uint32_t)out.size();
        e.addr = w.alloc(out.size());
        w.bytes(e.addr, out.data(), out.size());
        for (int k = 0; k < rank; k++) e.off.push_back(cidx[k] * d.chunk[k]);
        e.off.push_back(0);
        entries.push_back(std::move(e));
        for (int k = rank - 1; k >= 0; k--) {  // next chunk, last dimension fastest
            if (++cidx[k] < nchunks[k]) break;
            cidx[k] = 0;
        }
    }
    std::vector<uint64_t> end_off;  // the key after the last chunk
    for (int k = 0; k < rank; k++) end_off.push_back(k == 0 ? nchunks[0] * d.chunk[0] : 0);
    end_off.push_back(0);
    auto write_node = [&](int level, const std::vector<Entry> &es_, const std::vector<uint64_t> &last_key) {
        const uint64_t a = w.alloc(node_bytes);
        w.bytes(a, "TREE", 4);
        w.w8(a + 4, 1);
        w.w8(a + 5, level);
        w.w16(a + 6, (uint32_t)es_.size());
        w.w64(a + 8, UNDEF);
        w.w64(a + 16, UNDEF);
        uint64_t p = a + 24;
        for (const Entry &e : es_) {
            w.w32(p, e.size);
            w.w32(p + 4, 0);
            for (int k = 0; k <= rank; k++) w.w64(p + 8 + 8ull * k, e.off[k]);
            w.w64(p + keysz, e.addr);
            p += keysz + 8;
        }
        w.w32(p, 0);
        w.w32(p + 4, 0);
        for (int k = 0; k <= rank; k++) w.w64(p + 8 + 8ull * k, last_key[k]);
        return a;
    };
    if (entries.empty()) return UNDEF;
    if ((int)entries.size() <= 2 * K) return write_node(0, entries, end_off);
    if (entries.size() > (size_t)(2 * K) * (2 * K)) throw Error("too many chunks (at most 4096 per dataset)");
    std::vector<Entry> upper;
    for (size_t i = 0; i < entries.size(); i += 2 * K) {
        const size_t n = std::min<size_t>(2 * K, entries.size() - i);
        std::vector<Entry> leaf(entries.begin() + i, entries.begin() + i + n);
        const std::vector<uint64_t> &lk = i + n < entries.size() ? entries[i + n].off : end_off;
        Entry u;
        u.addr = write_node(0, leaf, lk);
        u.size = leaf[0].size;
        u.off = leaf[0].off;
        upper.push_back(std::move(u));
    }
    return write_node(1, upper, end_off);
}

uint64_t write_dataset(Wr &w, const Node &node)
{
    const Dataset &d = node.ds;
    const int rank = (int)d.dims.size();
    const bool chunked = !d.chunk.empty() && rank > 0 && d.count() > 0;
    if (chunked && (int)d.chunk.size() != rank) throw Error("chunk rank differs from the dataset rank");
    if (chunked)
        for (uint64_t c : d.chunk)
            if (c == 0) throw Error("chunk extent 0");
    const size_t es = d.type == Type::STR ? d.elem_size : type_size(d.type);
    const uint64_t nbytes = d.data.size();
    uint64_t daddr = UNDEF;
    if (chunked) daddr = write_chunks(w, d, es);
    else if (nbytes) {
        daddr = w.alloc(nbytes);
        w.bytes(daddr, d.data.data(), nbytes);
    }
    uint8_t tbody[24];
    const size_t tlen = datatype_body(d, tbody);
    const size_t space_sz = 8 + 8ull * rank;
    const size_t type_sz = (tlen + 7) & ~7ull;
    const int nfilters = chunked ? (d.shuffle ? 1 : 0) + (d.deflate_level > 0 ? 1 : 0) : 0;
    // pipeline v1: 8 bytes, then per filter id(2) name length(2) flags(2) n client values(2) + values (padded to 8)
    const size_t pipe_sz = nfilters ? 8 + (size_t)nfilters * 16 : 0;
    const size_t fill_sz = 8;
    const size_t layout_sz = chunked ? ((2 + 1 + 8 + 4ull * (rank + 1) + 7) & ~7ull) : 24;
    const int nmsg = 4 + (nfilters ? 1 : 0) + (int)node.extra.size();
    const size_t hsize = (size_t)(4 + (nfilters ? 1 : 0)) * 8 + space_sz + type_sz + fill_sz + layout_sz + pipe_sz + extra_bytes(node);
    const uint64_t o = w.alloc(16 + hsize);
    w.w8(o, 1);
    w.w16(o + 2, (uint32_t)nmsg);
    w.w32(o + 4, 1);
    w.w32(o + 8, (uint32_t)hsize);
    uint64_t p = o + 16;
    auto msg = [&](int type, size_t size, int flags) {
        w.w16(p, type);
        w.w16(p + 2, (uint32_t)size);
        w.w8(p + 4, flags);
        uint64_t body = p + 8;
        p += 8 + size;
        return body;
    };
    uint64_t q = msg(0x0001, space_sz, 0);  // dataspace v1
    w.w8(q, 1);
    w.w8(q + 1, rank);
    for (int i = 0; i < rank; i++) w.w64(q + 8 + 8ull * i, d.dims[i]);
    q = msg(0x0003, type_sz, 1);            // datatype v1
    w.bytes(q, tbody, tlen);
    q = msg(0x0005, fill_sz, 1);            // fill value v1: late allocation, write if set, defined, size 0
    w.w8(q, 1); w.w8(q + 1, chunked ? 3 : 2); w.w8(q + 2, 2); w.w8(q + 3, 1);
    if (nfilters) {
        q = msg(0x000B, pipe_sz, 0);        // filter pipeline v1
        w.w8(q, 1);
        w.w8(q + 1, nfilters);
        uint64_t f = q + 8;
        if (d.shuffle) {
            w.w16(f, 2); w.w16(f + 2, 0); w.w16(f + 4, 1); w.w16(f + 6, 1);   // shuffle: optional, one value = element size
            w.w32(f + 8, (uint32_t)es);
            f += 16;
        }
        if (d.deflate_level > 0) {
            w.w16(f, 1); w.w16(f + 2, 0); w.w16(f + 4, 1); w.w16(f + 6, 1);   // deflate: optional, one value = level
            w.w32(f + 8, (uint32_t)d.deflate_level);
            f += 16;
        }
    }
    q = msg(0x0008, layout_sz, 0);          // layout v3
    w.w8(q, 3);
    if (chunked) {
        w.w8(q + 1, 2);
        w.w8(q + 2, rank + 1);
        w.w64(q + 3, daddr);
        for (int i = 0; i < rank; i++) w.w32(q + 11 + 4ull * i, (uint32_t)d.chunk[i]);
        w.w32(q + 11 + 4ull * rank, (uint32_t)es);
    } else {
        w.w8(q + 1, 1);
        w.w64(q + 2, daddr);
        w.w64(q + 10, nbytes);
    }
    write_extra(w, p, node);
    return o;
}

struct GroupAddr { uint64_t ohdr, btree, heap; };

GroupAddr write_group(Wr &w, const Node &g)
{
    // children first
    std::vector<std::pair<std::string, uint64_t>> entries;  // std::map order == strcmp order for ASCII
    for (auto &kv : g.children) {
        uint64_t addr = kv.second->is_group ? write_group(w, *kv.second).ohdr : write_dataset(w, *kv.second);
        entries.emplace_back(kv.first, addr);
    }
    if ((int)entries.size() > 2 * kLeafK * 2 * kIntK) throw Error("too many entries in one group");
    // local heap: offset 0 = empty string, then the names, then one free block
    std::vector<uint64_t> name_off;
    uint64_t hsz = 8;
    for (auto &e : entries) {
        name_off.push_back(hsz);
        hsz += (e.first.size() + 1 + 7) & ~7ull;
    }
    const uint64_t free_off = hsz;
    hsz += 32;
    const uint64_t hdata = w.alloc(hsz);
    for (size_t i = 0; i < entries.size(); i++) w.bytes(hdata + name_off[i], entries[i].first.c_str(), entries[i].first.size());
    w.w64(hdata + free_off, 1);        // next free block: none
    w.w64(hdata + free_off + 8, 32);   // size of this free block
    const uint64_t heap = w.alloc(32);
    w.bytes(heap, "HEAP", 4);
    w.w64(heap + 8, hsz);
    w.w64(heap + 16, free_off);
    w.w64(heap + 24, hdata);
    // symbol nodes
    std::vector<uint64_t> snods, last_name;
    for (size_t i = 0; i < entries.size(); i += 2 * kLeafK) {
        const size_t n = std::min<size_t>(2 * kLeafK, entries.size() - i);
        const uint64_t s = w.alloc(8 + 40ull * 2 * kLeafK);
        w.bytes(s, "SNOD", 4);
        w.w8(s + 4, 1);
        w.w16(s + 6, (uint32_t)n);
        for (size_t k = 0; k < n; k++) {
            const uint64_t ep = s + 8 + 40 * k;
            w.w64(ep, name_off[i + k]);
            w.w64(ep + 8, entries[i + k].second);
        }
        snods.push_back(s);
        last_name.push_back(name_off[i + n - 1]);
    }
    // B-tree leaf node over the symbol nodes
    const uint64_t bt = w.alloc(24 + 8ull * (2 * kIntK + 1) + 8ull * 2 * kIntK);
    w.bytes(bt, "TREE", 4);
    w.w8(bt + 4, 0);
    w.w8(bt + 5, 0);
    w.w16(bt + 6, (uint32_t)snods.size());
    w.w64(bt + 8, UNDEF);
    w.w64(bt + 16, UNDEF);
    uint64_t p = bt + 24;
    w.w64(p, 0);  // key 0: the empty string
    p += 8;
    for (size_t i = 0; i < snods.size(); i++) {
        w.w64(p, snods[i]);
        w.w64(p + 8, last_name[i]);
        p += 16;
    }
    // object header: symbol-table message + a NIL message
    const uint64_t o = w.alloc(16 + 32 + extra_bytes(g));
    w.w8(o, 1);
    w.w16(o + 2, 2 + (uint32_t)g.extra.size());
    w.w32(o + 4, 1);
    w.w32(o + 8, 32 + (uint32_t)extra_bytes(g));
    w.w16(o + 16, 0x0011);
    w.w16(o + 18, 16);
    w.w64(o + 24, bt);
    w.w64(o + 32, heap);
    w.w16(o + 40, 0x0000);
    w.w16(o + 42, 0);
    uint64_t xp = o + 48;
    write_extra(w, xp, g);
    return GroupAddr{o, bt, heap};
}

}  // namespace

void File::save(const std::string &path) const
{
    Wr w;
    const uint64_t sb = w.alloc(96);
    GroupAddr rg = write_group(w, root);
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    w.bytes(sb, sig, 8);
    w.w8(sb + 13, 8);
    w.w8(sb + 14, 8);
    w.w16(sb + 16, kLeafK);
    w.w16(sb + 18, kIntK);
    w.w64(sb + 24, 0);            // base address
    w.w64(sb + 32, UNDEF);        // free-space info
    w.w64(sb + 48, UNDEF);        // driver info
    // root symbol table entry: cache type 1 = group, scratch = B-tree + heap addresses
    w.w64(sb + 56, 0);
    w.w64(sb + 64, rg.ohdr);
    w.w32(sb + 72, 1);
    w.w64(sb + 80, rg.btree);
    w.w64(sb + 88, rg.heap);
    const uint64_t eof = (w.b.size() + 7) & ~7ull;
    w.b.resize(eof, 0);
    w.w64(sb + 40, eof);          // end-of-file address
    const std::string tmp = path + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) throw Error("cannot write " + tmp);
    const bool ok = fwrite(w.b.data(), 1, w.b.size(), f) == w.b.size();
    fclose(f);
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) throw Error("cannot write " + path);
}

}  // namespace h5lite
