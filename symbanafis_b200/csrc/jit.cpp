// jit.cpp -- run-time compilation of specialised kernels (spec.h) with NVRTC, for sm_100a only.
//
// libnvrtc is loaded with dlopen on first use: the library itself does not link against it, program creation and the
// interpreter work without it, and saf_program_specialize reports SAF_ERR_UNSUPPORTED with the reason when it is
// missing.  The CUDA toolkit's copy is tried first (the build's own compiler version: the interpreter's cubin and the
// specialised kernels then come from the same front end and the same libdevice), then whatever the loader finds.
//
// The headers a generated kernel includes -- spec_prelude.cuh, devmath.cuh, leanmath.cuh, saf_isa.h -- are embedded in
// this library as text (.incbin) and handed to NVRTC as in-memory headers, together with two-line stand-ins for the
// <cstdint> / <cfloat> they include (NVRTC has no host headers).
#include "spec.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/saf_b200.h"

#define SAF_EMBED(sym, path)                                                                                \
    asm(".section .rodata\n.global " #sym "\n.balign 16\n" #sym ":\n.incbin \"" path "\"\n.byte 0\n.previous\n"); \
    extern "C" const char sym[];
SAF_EMBED(saf_text_spec_prelude, SAF_SRC_DIR "/spec_prelude.cuh")
SAF_EMBED(saf_text_devmath, SAF_SRC_DIR "/devmath.cuh")
SAF_EMBED(saf_text_leanmath, SAF_SRC_DIR "/leanmath.cuh")
SAF_EMBED(saf_text_saf_isa, SAF_SRC_DIR "/saf_isa.h")

namespace saf {

namespace {

const char kCstdint[] =
    "#pragma once\n"
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n"
    "typedef unsigned long long uintptr_t; typedef long long intptr_t;\n";
const char kCfloat[] =
    "#pragma once\n"
    "#define DBL_MAX 0x1.fffffffffffffp1023\n#define DBL_MIN 0x1p-1022\n#define DBL_EPSILON 0x1p-52\n";

typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    void *lib = nullptr;
    int (*Version)(int *, int *) = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram *) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string why_not;
    bool tried = false;
};
Nvrtc g_rtc;
std::mutex g_rtc_mu;

template <typename F>
bool sym(void *lib, const char *name, F *out) {
    *out = reinterpret_cast<F>(dlsym(lib, name));
    return *out != nullptr;
}

bool load_nvrtc() { // under g_rtc_mu
    if (g_rtc.tried) return g_rtc.lib != nullptr;
    g_rtc.tried = true;
    const char *override_path = std::getenv("SAF_NVRTC"); // full path of a libnvrtc to use
    const char *candidates[] = {override_path, "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so.13",
                                "libnvrtc.so"};
    void *lib = nullptr;
    for (const char *c : candidates) {
        if (c == nullptr || *c == 0) continue;
        lib = dlopen(c, RTLD_NOW | RTLD_LOCAL);
        if (lib != nullptr) break;
    }
    if (lib == nullptr) {
        g_rtc.why_not = "libnvrtc not found (tried SAF_NVRTC, /usr/local/cuda/lib64/libnvrtc.so.12, libnvrtc.so.12)";
        return false;
    }
    Nvrtc &r = g_rtc;
    const bool ok = sym(lib, "nvrtcVersion", &r.Version) && sym(lib, "nvrtcCreateProgram", &r.CreateProgram) &&
                    sym(lib, "nvrtcDestroyProgram", &r.DestroyProgram) && sym(lib, "nvrtcCompileProgram", &r.CompileProgram) &&
                    sym(lib, "nvrtcGetCUBINSize", &r.GetCUBINSize) && sym(lib, "nvrtcGetCUBIN", &r.GetCUBIN) &&
                    sym(lib, "nvrtcGetProgramLogSize", &r.GetProgramLogSize) && sym(lib, "nvrtcGetProgramLog", &r.GetProgramLog) &&
                    sym(lib, "nvrtcGetErrorString", &r.GetErrorString);
    if (!ok) {
        g_rtc.why_not = "libnvrtc lacks an entry point this library needs (nvrtcGetCUBIN needs CUDA 11.1 or later)";
        dlclose(lib);
        return false;
    }
    int major = 0, minor = 0;
    r.Version(&major, &minor);
    if (major < 12 || (major == 12 && minor < 8)) {
        g_rtc.why_not = "libnvrtc " + std::to_string(major) + "." + std::to_string(minor) + " cannot target sm_100a (12.8 or later needed)";
        dlclose(lib);
        return false;
    }
    r.lib = lib;
    return true;
}

} // namespace

bool jit_available(std::string &why_not) {
    std::lock_guard<std::mutex> lk(g_rtc_mu);
    const bool ok = load_nvrtc();
    if (!ok) why_not = g_rtc.why_not;
    return ok;
}

namespace {
// Optional disk cache of compiled kernels (SAF_SPEC_CACHE_DIR=<directory>, off when unset): a process that creates the
// same programs again -- a restarted service, a test-suite -- skips NVRTC.  The key covers everything that decides
// the cubin: the generated text, the embedded headers it includes, the NVRTC version, the target.
uint64_t fnv1a(const char *p, size_t n, uint64_t h) {
    for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)p[i]) * 0x100000001b3ull;
    return h;
}
std::string cache_path(const std::string &text, int major, int minor) {
    const char *dir = std::getenv("SAF_SPEC_CACHE_DIR");
    if (dir == nullptr || *dir == 0) return std::string();
    uint64_t h1 = 0xcbf29ce484222325ull, h2 = 0x84222325cbf29ce4ull;
    const char *parts[] = {text.c_str(), saf_text_spec_prelude, saf_text_devmath, saf_text_leanmath, saf_text_saf_isa};
    for (const char *p : parts) {
        const size_t n = std::strlen(p);
        h1 = fnv1a(p, n, h1);
        h2 = fnv1a(p, n, h2 ^ n);
    }
    char name[160];
    std::snprintf(name, sizeof name, "/saf_spec_sm100a_nvrtc%d.%d_%016llx%016llx.cubin", major, minor, (unsigned long long)h1,
                  (unsigned long long)h2);
    return std::string(dir) + name;
}
bool cache_read(const std::string &path, std::vector<char> &cubin) {
    if (path.empty()) return false;
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    bool ok = n > 64;
    if (ok) {
        cubin.resize((size_t)n);
        ok = std::fread(cubin.data(), 1, (size_t)n, f) == (size_t)n && std::memcmp(cubin.data(), "\177ELF", 4) == 0;
    }
    std::fclose(f);
    if (!ok) cubin.clear();
    return ok;
}
void cache_write(const std::string &path, const std::vector<char> &cubin) {
    if (path.empty()) return;
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f) return; // an unwritable cache directory is not an error
    const bool ok = std::fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    std::fclose(f);
    if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}
} // namespace

int jit_compile(const std::string &text, std::vector<char> &cubin, std::string &err) {
    {
        std::lock_guard<std::mutex> lk(g_rtc_mu);
        if (!load_nvrtc()) {
            err = g_rtc.why_not;
            return SAF_ERR_UNSUPPORTED;
        }
    }
    const Nvrtc &r = g_rtc; // read-only from here on; NVRTC itself may be called from several threads
    int v_major = 0, v_minor = 0;
    r.Version(&v_major, &v_minor);
    const std::string cached = cache_path(text, v_major, v_minor);
    if (cache_read(cached, cubin)) return SAF_OK;
    const char *headers[] = {saf_text_spec_prelude, saf_text_devmath, saf_text_leanmath, saf_text_saf_isa, kCstdint, kCstdint, kCfloat};
    const char *names[] = {"spec_prelude.cuh", "devmath.cuh", "leanmath.cuh", "saf_isa.h", "cstdint", "stdint.h", "cfloat"};
    nvrtcProgram prog = nullptr;
    int rc = r.CreateProgram(&prog, text.c_str(), "saf_spec.cu", 7, headers, names);
    if (rc != 0) {
        err = std::string("nvrtcCreateProgram: ") + r.GetErrorString(rc);
        return SAF_ERR_CUDA;
    }
    // -fmad stays at its default (true) exactly as in the build of kernel.cu: the CUDA math library needs it, and the
    // reference-mirroring arithmetic is written with explicit-rounding intrinsics that are never contracted.
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device", "-DSAF_NVRTC=1"};
    rc = r.CompileProgram(prog, 5, opts);
    if (rc != 0) {
        size_t n = 0;
        r.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) r.GetProgramLog(prog, &log[0]);
        if (log.size() > 4000) log.resize(4000);
        err = std::string("nvrtcCompileProgram: ") + r.GetErrorString(rc) + "\n" + log;
        r.DestroyProgram(&prog);
        return SAF_ERR_CUDA;
    }
    size_t n = 0;
    rc = r.GetCUBINSize(prog, &n);
    if (rc == 0 && n != 0) {
        cubin.resize(n);
        rc = r.GetCUBIN(prog, cubin.data());
    }
    r.DestroyProgram(&prog);
    if (rc != 0 || n == 0) {
        err = std::string("nvrtcGetCUBIN: ") + (rc ? r.GetErrorString(rc) : "empty image");
        return SAF_ERR_CUDA;
    }
    cache_write(cached, cubin);
    return SAF_OK;
}

} // namespace saf
