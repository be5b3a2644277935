// h5lite_c.cpp -- plain-C surface of h5lite for the ctypes layer (xpcs-eigen_b200/h5lite.py):
// tests and tools read and write the same HDF5 files the host `corr` does.
#include <cstring>
#include <string>

#include "h5lite.h"

using namespace h5lite;

static thread_local std::string g_err;

extern "C" {

const char *h5l_error() { return g_err.c_str(); }

void *h5l_new() { return new File(); }

void *h5l_open(const char *path)
{
    try {
        return new File(File::load(path));
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

int h5l_save(void *h, const char *path)
{
    try {
        ((File *)h)->save(path);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

void h5l_close(void *h) { delete (File *)h; }

// newline-separated child names of a group; returns the needed length
long h5l_list(void *h, const char *group, char *buf, long cap)
{
    std::string s;
    for (const std::string &n : ((File *)h)->list(group)) s += n + "\n";
    if (buf && cap > 0) {
        strncpy(buf, s.c_str(), (size_t)cap - 1);
        buf[cap - 1] = 0;
    }
    return (long)s.size() + 1;
}

// kind: 0 absent, 1 group, 2 dataset
int h5l_info(void *h, const char *path, int *type, int *rank, unsigned long long *dims, unsigned long long *elem_size,
             unsigned long long *nbytes)
{
    const Node *n = ((File *)h)->find(path);
    if (!n) return 0;
    if (n->is_group) return 1;
    *type = (int)n->ds.type;
    *rank = (int)n->ds.dims.size();
    for (size_t i = 0; i < n->ds.dims.size() && i < 8; i++) dims[i] = n->ds.dims[i];
    *elem_size = n->ds.elem_size;
    *nbytes = n->ds.data.size();
    return 2;
}

int h5l_read(void *h, const char *path, void *out, unsigned long long cap)
{
    const Node *n = ((File *)h)->find(path);
    if (!n || n->is_group) return -1;
    if (n->ds.data.size() > cap) return -2;
    if (!n->ds.data.empty()) memcpy(out, n->ds.data.data(), n->ds.data.size());
    return 0;
}

int h5l_put(void *h, const char *path, int type, int rank, const unsigned long long *dims, const void *data,
            unsigned long long str_len)
{
    try {
        std::vector<uint64_t> d(dims, dims + rank);
        ((File *)h)->put(path, (Type)type, d, data, (size_t)str_len);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// storage of a dataset at the next save(): rank chunk extents (rank 0 = contiguous), deflate level, byte shuffle
int h5l_set_storage(void *h, const char *path, int rank, const unsigned long long *chunk, int deflate_level, int shuffle)
{
    Node *n = ((File *)h)->find(path);
    if (!n || n->is_group) return -1;
    n->ds.chunk.assign(chunk, chunk + rank);
    n->ds.deflate_level = deflate_level;
    n->ds.shuffle = shuffle != 0;
    return 0;
}

// attribute / comment messages carried by the object at `path`: count, and raw body i (returns its length)
int h5l_extra_count(void *h, const char *path)
{
    const Node *n = ((File *)h)->find(path);
    return n ? (int)n->extra.size() : -1;
}

long h5l_extra_get(void *h, const char *path, int i, int *type, void *out, long cap)
{
    const Node *n = ((File *)h)->find(path);
    if (!n || i < 0 || i >= (int)n->extra.size()) return -1;
    const RawMessage &x = n->extra[(size_t)i];
    if (type) *type = x.type;
    if (out && cap >= (long)x.body.size() && !x.body.empty()) memcpy(out, x.body.data(), x.body.size());
    return (long)x.body.size();
}

// newline-separated: which = 0 lossy, 1 notes
long h5l_report(void *h, int which, char *buf, long cap)
{
    std::string s;
    for (const std::string &n : (which == 0 ? ((File *)h)->lossy : ((File *)h)->notes)) s += n + "\n";
    if (buf && cap > 0) {
        strncpy(buf, s.c_str(), (size_t)cap - 1);
        buf[cap - 1] = 0;
    }
    return (long)s.size() + 1;
}

}  // extern "C"
