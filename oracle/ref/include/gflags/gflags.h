/*
 * gflags/gflags.h -- a 60-line stand-in for the gflags library (a git submodule of the
 * reference that needs cmake-generated headers), covering exactly what the reference's
 * main.cpp:86-110 uses: DEFINE_bool / DEFINE_int32 / DEFINE_string and
 * gflags::ParseCommandLineFlags(&argc, &argv, true) with --flag, --flag=value, --noflag.
 * TEST INFRASTRUCTURE (oracle/); written from the gflags documentation, no gflags code copied.
 */
#ifndef XPCS_ORACLE_GFLAGS_SHIM_H
#define XPCS_ORACLE_GFLAGS_SHIM_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

namespace gflags {
struct FlagReg {
    const char *name;
    int kind;  // 0 bool, 1 int32, 2 string
    void *ptr;
};
inline std::vector<FlagReg> &registry()
{
    static std::vector<FlagReg> r;
    return r;
}
struct Registrar {
    Registrar(const char *n, int k, void *p) { registry().push_back(FlagReg{n, k, p}); }
};
inline bool set_flag(const std::string &name, const char *value)
{
    for (auto &f : registry()) {
        if (name == f.name) {
            if (f.kind == 0) *(bool *)f.ptr = !value || !strcmp(value, "true") || !strcmp(value, "1");
            else if (f.kind == 1) *(int32_t *)f.ptr = value ? atoi(value) : 0;
            else *(std::string *)f.ptr = value ? value : "";
            return true;
        }
        if (f.kind == 0 && name == std::string("no") + f.name) {
            *(bool *)f.ptr = false;
            return true;
        }
    }
    return false;
}
inline uint32_t ParseCommandLineFlags(int *argc, char ***argv, bool remove_flags)
{
    int out = 1;
    for (int i = 1; i < *argc; i++) {
        char *a = (*argv)[i];
        if (a[0] == '-' && a[1] != '\0') {
            const char *p = a + 1;
            if (*p == '-') p++;
            std::string s(p);
            size_t eq = s.find('=');
            std::string name = eq == std::string::npos ? s : s.substr(0, eq);
            const char *val = eq == std::string::npos ? nullptr : p + eq + 1;
            bool needs_value = false;
            for (auto &f : registry())
                if (name == f.name && f.kind != 0 && !val) needs_value = true;
            if (needs_value && i + 1 < *argc) val = (*argv)[++i];
            set_flag(name, val);
            if (remove_flags) continue;
        }
        (*argv)[out++] = a;
    }
    if (remove_flags) *argc = out;
    return (uint32_t)out;
}
}  // namespace gflags

#define DEFINE_bool(name, dflt, help) \
    bool FLAGS_##name = dflt;          \
    static gflags::Registrar reg_##name(#name, 0, &FLAGS_##name)
#define DEFINE_int32(name, dflt, help) \
    int32_t FLAGS_##name = dflt;        \
    static gflags::Registrar reg_##name(#name, 1, &FLAGS_##name)
#define DEFINE_string(name, dflt, help) \
    std::string FLAGS_##name = dflt;     \
    static gflags::Registrar reg_##name(#name, 2, &FLAGS_##name)
#endif
