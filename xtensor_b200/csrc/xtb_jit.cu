// xtb_jit.cu -- run-time specialisation of the hand-written kernel templates.
//
// xtensor fuses arbitrary expression trees; only a handful of them can be instantiated ahead
// of time (xtb_static_programs.cuh).  For every other program the SAME templates (xtb_ew.cuh,
// xtb_reduce.cuh) are instantiated at run time with the program as a template constant: the
// instruction list is printed into a tiny translation unit, NVRTC compiles it for the device's
// architecture (~2 s, once per distinct program and kernel shape, cached for the process), and
// the kernel is launched through the driver API.  Nothing is traced or generated beyond the
// constexpr program table -- the device code is the code in this directory.
// NVRTC / libcuda are dlopen'ed; if they are unavailable (or XTB_NO_JIT is set) the caller falls
// back to the interpreter kernels, still on the GPU.
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>
#include "xtb_common.hpp"
#include "xtb_jit.hpp"

namespace xtb {

namespace {

typedef int (*fn_nvrtcCreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*);
typedef int (*fn_nvrtcCompileProgram)(void*, int, const char* const*);
typedef int (*fn_nvrtcGetSize)(void*, size_t*);
typedef int (*fn_nvrtcGetData)(void*, char*);
typedef int (*fn_nvrtcDestroyProgram)(void**);
typedef int (*fn_nvrtcAddNameExpression)(void*, const char*);
typedef int (*fn_nvrtcGetLoweredName)(void*, const char*, const char**);
typedef int (*fn_cuModuleLoadData)(void**, const void*);
typedef int (*fn_cuModuleGetFunction)(void**, void*, const char*);
typedef int (*fn_cuLaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**);
typedef int (*fn_cuFuncSetAttribute)(void*, int, int);

struct Api {
    bool tried = false, ok = false;
    fn_nvrtcCreateProgram create = nullptr;
    fn_nvrtcCompileProgram compile = nullptr;
    fn_nvrtcGetSize cubin_size = nullptr, log_size = nullptr;
    fn_nvrtcGetData cubin = nullptr, log = nullptr;
    fn_nvrtcDestroyProgram destroy = nullptr;
    fn_nvrtcAddNameExpression add_name = nullptr;
    fn_nvrtcGetLoweredName lowered = nullptr;
    fn_cuModuleLoadData module_load = nullptr;
    fn_cuModuleGetFunction module_get = nullptr;
    fn_cuLaunchKernel launch = nullptr;
    fn_cuFuncSetAttribute func_attr = nullptr;
    std::string csrc_dir, cuda_inc, arch;
};
Api g_api;
std::mutex g_mutex;
std::map<std::string, void*> g_cache;
int g_compiles = 0;

void* sym(void* h, const char* name) { return h ? dlsym(h, name) : nullptr; }

bool load_api(const DeviceCtx* ctx) {
    if (g_api.tried) return g_api.ok;
    g_api.tried = true;
    void* rtc = nullptr;
    const char* rtc_names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : rtc_names) {
        rtc = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (rtc) break;
    }
    void* cu = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!rtc || !cu) return false;
    g_api.create = (fn_nvrtcCreateProgram) sym(rtc, "nvrtcCreateProgram");
    g_api.compile = (fn_nvrtcCompileProgram) sym(rtc, "nvrtcCompileProgram");
    g_api.cubin_size = (fn_nvrtcGetSize) sym(rtc, "nvrtcGetCUBINSize");
    g_api.cubin = (fn_nvrtcGetData) sym(rtc, "nvrtcGetCUBIN");
    g_api.log_size = (fn_nvrtcGetSize) sym(rtc, "nvrtcGetProgramLogSize");
    g_api.log = (fn_nvrtcGetData) sym(rtc, "nvrtcGetProgramLog");
    g_api.destroy = (fn_nvrtcDestroyProgram) sym(rtc, "nvrtcDestroyProgram");
    g_api.add_name = (fn_nvrtcAddNameExpression) sym(rtc, "nvrtcAddNameExpression");
    g_api.lowered = (fn_nvrtcGetLoweredName) sym(rtc, "nvrtcGetLoweredName");
    g_api.module_load = (fn_cuModuleLoadData) sym(cu, "cuModuleLoadData");
    g_api.module_get = (fn_cuModuleGetFunction) sym(cu, "cuModuleGetFunction");
    g_api.launch = (fn_cuLaunchKernel) sym(cu, "cuLaunchKernel");
    g_api.func_attr = (fn_cuFuncSetAttribute) sym(cu, "cuFuncSetAttribute");
    if (!g_api.create || !g_api.compile || !g_api.cubin_size || !g_api.cubin || !g_api.destroy || !g_api.add_name ||
        !g_api.lowered || !g_api.module_load || !g_api.module_get || !g_api.launch)
        return false;
    // the kernel sources sit next to the library: <repo>/xtensor_b200/lib/libxtb200.so -> ../csrc
    Dl_info info;
    if (!dladdr((void*) &load_api, &info) || !info.dli_fname) return false;
    std::string lib = info.dli_fname;
    const size_t slash = lib.rfind('/');
    g_api.csrc_dir = (slash == std::string::npos ? std::string(".") : lib.substr(0, slash)) + "/../csrc";
    if (const char* e = getenv("XTB_CSRC_DIR")) g_api.csrc_dir = e;
    const char* home = getenv("CUDA_HOME");
    g_api.cuda_inc = std::string(home ? home : "/usr/local/cuda") + "/include";
    FILE* f = fopen((g_api.csrc_dir + "/xtb_ew.cuh").c_str(), "r");
    if (!f) return false;
    fclose(f);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) return false;
    char arch[32];
    snprintf(arch, sizeof(arch), "sm_%d%d%s", prop.major, prop.minor, prop.major >= 9 ? "a" : "");
    g_api.arch = arch;
    g_api.ok = true;
    return true;
}

std::string program_literal(const xtb_program* p) {
    std::ostringstream o;
    o << "sprogs::SP{" << p->n_insns << ", {";
    for (int i = 0; i < p->n_insns; ++i) {
        const xtb_insn in = p->insns[i];
        o << "{" << (int) in.op << "," << (int) in.type << "," << (int) in.src << "," << (int) in.arg << "},";
    }
    o << "}, " << p->n_leaves << ", " << p->n_imms << "}";
    return o.str();
}

}  // namespace

bool jit_program_ok(const xtb_program* p) { return p->n_insns <= XTB_MAX_INSNS; }

// A compile costs ~2 s: only worth it when the problem is large (the interpreter kernels finish a
// small problem in microseconds).  XTB_JIT_MIN_ELEMS overrides the threshold (tests use 0).
bool jit_worthwhile(int64_t elements) {
    return elements >= (int64_t) options().jit_min_elems;
}

int jit_get(const DeviceCtx* ctx, const xtb_program* prog, const JitSpec& spec, void** fn) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (options().no_jit || !load_api(ctx)) return XTB_ERR_UNSUPPORTED;
    std::string key((const char*) prog->insns, sizeof(xtb_insn) * prog->n_insns);
    char tail[96];
    // the CUfunction belongs to the primary context of the device it was loaded on: the device is part of the key
    snprintf(tail, sizeof(tail), "|%d|%d|%d|%d|%d|%d|%d|dev%d", spec.kind, spec.w64, spec.V, spec.nd, spec.binop, spec.acc_rt, prog->n_leaves,
             ctx->device);
    key += tail;
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
        *fn = it->second;
        return it->second ? XTB_OK : XTB_ERR_UNSUPPORTED;
    }
    g_cache[key] = nullptr;  // a failed compile is not retried
    const char* S = spec.w64 ? "uint64_t" : "uint32_t";
    std::ostringstream src, name;
    src << "#include \"xtb_reduce.cuh\"\n"
        << "namespace xtb { namespace jit {\n"
        << "struct Tbl { static constexpr sprogs::SP progs[] = { " << program_literal(prog) << " }; };\n"
        << "} }\n";
    const std::string eval = "xtb::StaticEval<xtb::jit::Tbl, 0>";
    std::ostringstream acc;
    acc << "xtb::StaticAcc<" << spec.binop << ", " << spec.acc_rt << ">";
    switch (spec.kind) {
        case JIT_EW: name << "xtb::k_ew<" << eval << ", " << S << ", " << spec.V << ", " << spec.nd << ", xtb::kEwItems>"; break;
        case JIT_TILE: name << "xtb::k_ew_tile_static<" << eval << ", " << S << ">"; break;
        case JIT_RED_OUTER: name << "xtb::k_reduce_outer<" << eval << ", " << acc.str() << ", " << S << ", " << spec.V << ">"; break;
        case JIT_RED_INNER_WARP: name << "xtb::k_reduce_inner_warp<" << eval << ", " << acc.str() << ", " << S << ", " << spec.V << ">"; break;
        case JIT_RED_INNER_BLOCK: name << "xtb::k_reduce_inner_block<" << eval << ", " << acc.str() << ", " << S << ", " << spec.V << ">"; break;
        case JIT_RED_ROWS_EXACT: name << "xtb::k_reduce_rows_exact<" << eval << ", " << acc.str() << ", " << S << ", " << spec.V << ">"; break;
        default: return set_error(XTB_ERR_INVALID, "bad jit kind");
    }
    void* prog_h = nullptr;
    const std::string source = src.str(), expr = name.str();
    if (g_api.create(&prog_h, source.c_str(), "xtb_jit.cu", 0, nullptr, nullptr) != 0) return XTB_ERR_UNSUPPORTED;
    g_api.add_name(prog_h, expr.c_str());
    std::string dev_arch = g_api.arch;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess) {
            char a[32];
            snprintf(a, sizeof(a), "sm_%d%d%s", prop.major, prop.minor, prop.major >= 9 ? "a" : "");
            dev_arch = a;
        }
    }
    const std::string inc1 = "-I" + g_api.csrc_dir, inc2 = "-I" + g_api.cuda_inc, arch = "--gpu-architecture=" + dev_arch;
    const char* opts[] = {arch.c_str(), "--std=c++20", inc1.c_str(), inc2.c_str(), "--fmad=false", "-default-device"};
    const int rc = g_api.compile(prog_h, 6, opts);
    ++g_compiles;
    if (rc != 0) {
        size_t n = 0;
        std::string log;
        if (g_api.log_size && g_api.log && g_api.log_size(prog_h, &n) == 0 && n > 1) {
            log.resize(n);
            g_api.log(prog_h, log.data());
        }
        if (options().jit_verbose) fprintf(stderr, "[xtb jit] compile failed for %s:\n%s\n", expr.c_str(), log.c_str());
        g_api.destroy(&prog_h);
        return set_error(XTB_ERR_UNSUPPORTED, "jit compile failed: %.300s", log.c_str());
    }
    const char* lowered = nullptr;
    size_t csz = 0;
    if (g_api.lowered(prog_h, expr.c_str(), &lowered) != 0 || !lowered || g_api.cubin_size(prog_h, &csz) != 0 || csz == 0) {
        g_api.destroy(&prog_h);
        return set_error(XTB_ERR_UNSUPPORTED, "jit: no cubin");
    }
    std::vector<char> cubin(csz);
    g_api.cubin(prog_h, cubin.data());
    void* module = nullptr;
    void* func = nullptr;
    cudaFree(0);  // make sure the primary context is current for the driver API
    if (g_api.module_load(&module, cubin.data()) != 0 || g_api.module_get(&func, module, lowered) != 0) {
        g_api.destroy(&prog_h);
        return set_error(XTB_ERR_UNSUPPORTED, "jit: module load failed");
    }
    g_api.destroy(&prog_h);
    if (options().jit_verbose) fprintf(stderr, "[xtb jit] compiled %s\n", expr.c_str());
    g_cache[key] = func;
    *fn = func;
    return XTB_OK;
}

int jit_launch(void* fn, unsigned gx, unsigned gy, unsigned block, size_t smem, cudaStream_t stream, const void* params) {
    void* args[1] = {const_cast<void*>(params)};
    const int rc = g_api.launch(fn, gx, gy, 1, block, 1, 1, (unsigned) smem, (void*) stream, args, nullptr);
    if (rc != 0) return set_error(XTB_ERR_CUDA, "jit kernel launch failed (CUresult %d)", rc);
    return XTB_OK;
}

int jit_compile_count() { return g_compiles; }

}  // namespace xtb
