// lane_jit.cu -- run-time specialisation of the lane-per-problem kernel (fused_lane_dev.cuh) for layouts that have no
// compile-time instantiation in fused_lane.cu: the kernel template is instantiated for the caller's (n, orthant rows,
// cone count, cone dimension) with NVRTC the first time a handle with that layout wants it, and cached per process
// and device.  This is the GPU analogue of what the reference gets from Julia: `solve_socp` is specialised on the cone
// tuple's type parameters (src/Socp.jl:8-18, `@unroll` over cones) the first time it is called with a new layout.
// libnvrtc and libcuda are loaded lazily with dlopen, so the library has no link-time dependency on either; when they
// are missing, or the compilation fails, the caller falls back to fused_v2's one-warp teams.
#include "fused_lane.cuh"
#include "fused_v3.cuh"
#include "fused_v2.cuh"
#include <dlfcn.h>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace socp {
namespace {

typedef int nvrtcResult_t;
typedef struct _nvrtcProgram* nvrtcProgram_t;
typedef int CUresult_t;
typedef struct CUmod_st* CUmodule_t;
typedef struct CUfunc_st* CUfunction_t;

struct Api {
    bool ok = false;
    nvrtcResult_t (*CreateProgram)(nvrtcProgram_t*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult_t (*CompileProgram)(nvrtcProgram_t, int, const char* const*) = nullptr;
    nvrtcResult_t (*AddNameExpression)(nvrtcProgram_t, const char*) = nullptr;
    nvrtcResult_t (*GetLoweredName)(nvrtcProgram_t, const char*, const char**) = nullptr;
    nvrtcResult_t (*GetCUBINSize)(nvrtcProgram_t, size_t*) = nullptr;
    nvrtcResult_t (*GetCUBIN)(nvrtcProgram_t, char*) = nullptr;
    nvrtcResult_t (*GetProgramLogSize)(nvrtcProgram_t, size_t*) = nullptr;
    nvrtcResult_t (*GetProgramLog)(nvrtcProgram_t, char*) = nullptr;
    nvrtcResult_t (*DestroyProgram)(nvrtcProgram_t*) = nullptr;
    CUresult_t (*ModuleLoadData)(CUmodule_t*, const void*) = nullptr;
    CUresult_t (*ModuleGetFunction)(CUfunction_t*, CUmodule_t, const char*) = nullptr;
    CUresult_t (*FuncSetAttribute)(CUfunction_t, int, int) = nullptr;
    CUresult_t (*LaunchKernel)(CUfunction_t, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
};

template <class F>
bool sym(void* lib, const char* name, F& f) {
    f = reinterpret_cast<F>(dlsym(lib, name));
    return f != nullptr;
}

Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* rtc = nullptr;
        for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"})
            if ((rtc = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        void* cu = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!rtc || !cu) return;
        a.ok = sym(rtc, "nvrtcCreateProgram", a.CreateProgram) && sym(rtc, "nvrtcCompileProgram", a.CompileProgram) &&
               sym(rtc, "nvrtcAddNameExpression", a.AddNameExpression) && sym(rtc, "nvrtcGetLoweredName", a.GetLoweredName) &&
               sym(rtc, "nvrtcGetCUBINSize", a.GetCUBINSize) && sym(rtc, "nvrtcGetCUBIN", a.GetCUBIN) &&
               sym(rtc, "nvrtcGetProgramLogSize", a.GetProgramLogSize) && sym(rtc, "nvrtcGetProgramLog", a.GetProgramLog) &&
               sym(rtc, "nvrtcDestroyProgram", a.DestroyProgram) && sym(cu, "cuModuleLoadData", a.ModuleLoadData) &&
               sym(cu, "cuModuleGetFunction", a.ModuleGetFunction) && sym(cu, "cuFuncSetAttribute", a.FuncSetAttribute) &&
               sym(cu, "cuLaunchKernel", a.LaunchKernel);
    });
    return a;
}

// directory of the kernel sources: next to the library (.../lib/libsocp_b200.so -> .../csrc), or SOCP_B200_CSRC
std::string csrc_dir() {
    if (const char* e = getenv("SOCP_B200_CSRC")) return e;
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&csrc_dir), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t s1 = p.rfind('/');
        if (s1 != std::string::npos) {
            const size_t s2 = p.rfind('/', s1 - 1);
            if (s2 != std::string::npos) return p.substr(0, s2) + "/csrc";
        }
    }
    return "csrc";
}

using Key = std::pair<int, std::string>;      // device, the kernel instance
std::map<Key, CUfunction_t>& cache() { static std::map<Key, CUfunction_t> c; return c; }
std::mutex& cache_mutex() { static std::mutex m; return m; }

}  // namespace

// Compiles `inst` (an instantiation of a kernel template defined in `header`) once per process and device and returns
// its CUfunction with `smem` bytes of dynamic shared memory enabled; nullptr when NVRTC / the driver API are not
// available or the compilation fails (the reason goes to stderr when SOCP_B200_JIT_VERBOSE is set).
static void* jit_kernel(const char* header, const std::string& inst, int device, size_t smem) {
    Api& a = api();
    if (!a.ok) return nullptr;
    const Key key{device, inst};
    std::lock_guard<std::mutex> lock(cache_mutex());
    auto it = cache().find(key);
    if (it != cache().end()) return it->second;
    CUfunction_t fn = nullptr;
    const bool verbose = getenv("SOCP_B200_JIT_VERBOSE") != nullptr;
    const std::string src = std::string("#include \"") + header + "\"\n";
    nvrtcProgram_t prog = nullptr;
    if (a.CreateProgram(&prog, src.c_str(), "socp_b200_jit_instance.cu", 0, nullptr, nullptr) == 0) {
        const std::string inc = "-I" + csrc_dir();
        const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-diag-suppress=607", inc.c_str()};
        const char* lowered = nullptr;
        std::vector<char> cubin;
        if (a.AddNameExpression(prog, inst.c_str()) == 0 && a.CompileProgram(prog, 5, opts) == 0 &&
            a.GetLoweredName(prog, inst.c_str(), &lowered) == 0 && lowered) {
            size_t sz = 0;
            if (a.GetCUBINSize(prog, &sz) == 0 && sz > 0) {
                cubin.resize(sz);
                CUmodule_t mod = nullptr;
                cudaFree(0);                                     // the primary context of the current device
                if (a.GetCUBIN(prog, cubin.data()) == 0 && a.ModuleLoadData(&mod, cubin.data()) == 0 &&
                    a.ModuleGetFunction(&fn, mod, lowered) == 0 && fn) {
                    // CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8
                    if (a.FuncSetAttribute(fn, 8, (int)smem) != 0) fn = nullptr;
                }
            }
        } else if (verbose) {
            size_t n = 0;
            a.GetProgramLogSize(prog, &n);
            std::vector<char> log(n + 1, 0);
            a.GetProgramLog(prog, log.data());
            fprintf(stderr, "socp_b200: NVRTC could not build %s:\n%s\n", inst.c_str(), log.data());
        }
        a.DestroyProgram(&prog);
    }
    if (verbose) fprintf(stderr, "socp_b200: %s %s\n", inst.c_str(), fn ? "compiled at run time" : "NOT available");
    cache()[key] = fn;
    return fn;
}

// the lane-per-problem kernel for the plan's layout (shape 100)
void* lane_jit_get(const FLPlan& P, int device) {
    std::string groups;
    for (const auto& g : P.jgroups) groups += ", socp::ConeGroup<" + std::to_string(g.first) + ", " + std::to_string(g.second) + ">";
    const std::string inst = "socp::k_fused_lane<socp::LaneDimsG<" + std::to_string(P.jn) + ", " + std::to_string(P.jkpoc) + ", " +
                             std::to_string(P.jrs) + groups + ">, 32, " + std::to_string(P.pps / 32) + ">";
    return jit_kernel("fused_lane_dev.cuh", inst, device, P.smem);
}

// The whole-solve kernel of fused_v3.cuh with the layout as compile-time constants (Dims3Static): what fused3.cu
// instantiates for BASELINE.json's C2 only, for any layout with one orthant block and equal second-order cones whose
// singleton rows are the identity block or absent.  2.3x over the runtime-dimension instantiation on C2.
bool fused3_jit_shape(const F3Plan& P) {
    if (!P.fits || P.nw != 4 || P.nsoc < 1) return false;
    if (!(P.nsing == 0 || (P.ident && P.nsing == P.n))) return false;
    for (int i = 0; i < P.nsoc; ++i)
        if (P.soc_dim[i] != P.soc_dim[0] || P.soc_offs[i] != P.kpoc + i * P.soc_dim[0]) return false;
    return P.k == P.kpoc + P.nsoc * P.soc_dim[0];
}
void* fused3_jit_get(const F3Plan& P, int teams, int device) {
    if (!fused3_jit_shape(P)) return nullptr;
    const int maxt = P.nb <= 4 ? 3 : (P.nb <= 7 ? 7 : 9);
    const int minb = teams == 4 ? 1 : (P.nb <= 7 ? 4 : 3);
    char inst[320];
    snprintf(inst, sizeof inst, "socp::k_fused3<4, %d, %d, %d, socp::Dims3Static<4, %d, %d, %d, %d, %d, %d, %d, %d>>", teams, maxt,
             minb, P.n, P.p, P.kpoc, P.nsoc, P.soc_dim[0], P.d0, P.kd, P.ident ? 1 : 0);
    return jit_kernel("fused_v3.cuh", inst, device, P.smem * teams);
}
// fused_v2's kernel (one-warp teams for n <= 16; what fused_v3 declines) with the layout as compile-time constants
void* fused2_jit_get(const F2Plan& P) {
    if (!P.fits || P.nsoc < 1) return nullptr;
    if (DimsC2::matches(P) || DimsC3::matches(P)) return nullptr;       // instantiated at compile time
    for (int i = 0; i < P.nsoc; ++i)
        if (P.soc_dim[i] != P.soc_dim[0] || P.soc_offs[i] != P.kpoc + i * P.soc_dim[0]) return nullptr;
    if (P.k != P.kpoc + P.nsoc * P.soc_dim[0]) return nullptr;
    static const int cfg[4][3] = {{1, 3, 16}, {4, 3, 4}, {8, 4, 2}, {8, 5, 2}};      // NW, MAXT, MINB of fused2_launch per variant
    const int v = P.variant < 0 || P.variant > 3 ? 3 : P.variant;
    char inst[256];
    snprintf(inst, sizeof inst, "socp::k_fused2<%d, %d, %d, socp::DimsStatic<%d, %d, %d, %d, %d, %d>>", cfg[v][0], cfg[v][1],
             cfg[v][2], cfg[v][0], P.n, P.p, P.kpoc, P.nsoc, P.soc_dim[0]);
    return jit_kernel("fused_v2.cuh", inst, P.device, P.smem);
}
bool fused2_jit_launch(void* fn, const F2Plan& plan, const F2Args& args, int grid, cudaStream_t stream) {
    Api& a = api();
    if (!a.ok || !fn) return false;
    F2Args copy = args;
    void* params[] = {&copy};
    return a.LaunchKernel(reinterpret_cast<CUfunction_t>(fn), (unsigned)grid, 1, 1, (unsigned)(plan.nw * 32), 1, 1,
                          (unsigned)plan.smem, stream, params, nullptr) == 0;
}

bool fused3_jit_launch(void* fn, const F3Plan& plan, const F3Args& args, int teams, int grid, cudaStream_t stream) {
    Api& a = api();
    if (!a.ok || !fn) return false;
    F3Args copy = args;
    void* params[] = {&copy};
    return a.LaunchKernel(reinterpret_cast<CUfunction_t>(fn), (unsigned)grid, 1, 1, (unsigned)(128 * teams), 1, 1,
                          (unsigned)(plan.smem * teams), stream, params, nullptr) == 0;
}

bool lane_jit_launch(void* fn, const FLPlan& plan, FLArgs args, cudaStream_t stream) {
    Api& a = api();
    if (!a.ok || !fn) return false;
    int grid;
    fl_grid(plan, args.batch, 32, grid, args.cap);
    void* params[] = {&args};
    return a.LaunchKernel(reinterpret_cast<CUfunction_t>(fn), (unsigned)grid, 1, 1, (unsigned)(plan.pps / 32) * 32, 1, 1,
                          (unsigned)plan.smem, stream, params, nullptr) == 0;
}

}  // namespace socp
