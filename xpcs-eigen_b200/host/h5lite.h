// h5lite.h -- a small from-scratch reader/writer for the subset of the HDF5 FILE FORMAT the
// XPCS configuration / result files use.  The image (and the GPU box) has no libhdf5, and the
// reference keeps its whole user contract in HDF5 (Configuration::init, configuration.cpp:80-242;
// H5Result::write*, h5_result.cpp:56-347), so the host program carries its own implementation.
// Written from the format description in SURVEY.md Appendix D (decoded from a genuine file);
// no libhdf5 or reference code is used.
//
// Read:  superblock v0/v1, v1 object headers (with continuation blocks), symbol-table groups
//        (v1 B-tree + local heap + SNOD), dataspace v1/v2, datatypes fixed-point / float /
//        fixed string / variable-length string (global heap), layouts v1-v3 compact, contiguous
//        and chunked (v1 chunk B-tree, deflate + shuffle filters).
// Write: superblock v0, v1 object headers, symbol-table groups, contiguous datasets of the
//        native numeric types and fixed strings -- the oldest, simplest encodings every HDF5
//        library reads.  A file is held as an in-memory tree; save() rewrites it whole, which
//        gives the reference's "open RDWR, add or overwrite datasets" behaviour
//        (h5_result.cpp:67-103) without in-place B-tree surgery.
// Attribute and object-comment messages of groups and datasets are carried over VERBATIM (the reference opens
// the user's file read-write and only adds datasets, h5_result.cpp:67-103, so a rewrite must not strip
// metadata); an attribute whose bytes point elsewhere in the file (variable-length or reference data, shared
// datatypes) cannot be moved that way: it is listed in File::lossy and the host program then keeps a backup of
// the input.  Datasets found chunked / compressed / with variable-length strings are rewritten contiguous /
// fixed-length (same values; listed in File::notes).
// Not supported (reported as errors, never silently skipped): superblock v2/v3, v2 object
// headers ("OHDR"), compound / array / reference datatypes, external storage.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace h5lite {

enum class Type { I8, U8, I16, U16, I32, U32, I64, U64, F32, F64, STR };

size_t type_size(Type t);

struct Dataset {
    Type type = Type::F32;
    size_t elem_size = 4;            // for STR: the fixed string length
    std::vector<uint64_t> dims;      // empty = scalar
    std::vector<uint8_t> data;       // little-endian, row-major
    // storage chosen by save(): empty chunk = contiguous; else chunked (one chunk extent per dimension, edge chunks
    // padded), each chunk optionally byte-shuffled and deflated -- what H5Pset_chunk / H5Pset_shuffle /
    // H5Pset_deflate produce (the reference stores C2T_all/g2_* as one deflate-6 chunk, corr.cpp:883-923)
    std::vector<uint64_t> chunk;
    int deflate_level = 0;           // 0 = not compressed
    bool shuffle = false;
    uint64_t count() const;
    // conversions (numeric types convert like H5Dread with a native memory type)
    std::vector<int32_t> as_i32() const;
    std::vector<int64_t> as_i64() const;
    std::vector<float> as_f32() const;
    std::vector<double> as_f64() const;
    std::string as_string() const;   // STR datasets (first element)
    double scalar() const;
};

struct RawMessage {   // an object-header message kept as found (attribute 0x000C, comment 0x000D)
    uint16_t type = 0;
    uint8_t flags = 0;
    std::vector<uint8_t> body;   // length is a multiple of 8 (version-1 object headers)
};

struct Node {
    bool is_group = true;
    std::vector<RawMessage> extra;                           // carried over by save()
    std::map<std::string, std::unique_ptr<Node>> children;  // groups (sorted by name, as SNODs are)
    Dataset ds;                                              // datasets
};

class Error : public std::runtime_error {
public:
    explicit Error(const std::string &m) : std::runtime_error("h5lite: " + m) {}
};

class File {
public:
    File();
    static File load(const std::string &path);   // throws Error
    void save(const std::string &path) const;    // throws Error
    Node *find(const std::string &path);         // nullptr when absent
    const Node *find(const std::string &path) const;
    bool has(const std::string &path) const { return find(path) != nullptr; }
    const Dataset &dataset(const std::string &path) const;  // throws when absent or a group
    Node &make_group(const std::string &path);               // creates intermediate groups
    Dataset &put(const std::string &path, Type t, const std::vector<uint64_t> &dims, const void *data,
                 size_t str_len = 0);                        // create or overwrite
    Dataset &put_string(const std::string &path, const std::string &value);
    std::vector<std::string> list(const std::string &group) const;
    Node root;
    std::vector<std::string> lossy;   // content of the loaded file a save() cannot reproduce
    std::vector<std::string> notes;   // content a save() reproduces in another encoding (same values)
};

}  // namespace h5lite
