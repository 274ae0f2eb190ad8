// Run-time compilation of generated Gauss-point kernels (the role FFC + dijitso
// play behind dolfin.assemble in the reference, common.py:1215-1216): CUDA C
// source -> NVRTC -> sm_100a cubin -> driver-API module -> launch on the
// caller's stream.  libnvrtc and libcuda are opened lazily with dlopen so that
// the library still loads (and exports every symbol) on a box without a GPU.
#include "tg_common.cuh"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

typedef int nvrtcResult_t;
typedef struct _nvrtcProgram* nvrtcProgram_t;
typedef int CUresult_t;
typedef struct CUmod_st* CUmodule_t;
typedef struct CUfunc_st* CUfunction_t;

struct TgNvrtc {
  void* h;
  nvrtcResult_t (*CreateProgram)(nvrtcProgram_t*, const char*, const char*, int, const char* const*,
                                 const char* const*);
  nvrtcResult_t (*CompileProgram)(nvrtcProgram_t, int, const char* const*);
  nvrtcResult_t (*GetProgramLogSize)(nvrtcProgram_t, size_t*);
  nvrtcResult_t (*GetProgramLog)(nvrtcProgram_t, char*);
  nvrtcResult_t (*GetCUBINSize)(nvrtcProgram_t, size_t*);
  nvrtcResult_t (*GetCUBIN)(nvrtcProgram_t, char*);
  nvrtcResult_t (*DestroyProgram)(nvrtcProgram_t*);
};
struct TgCuda {
  void* h;
  CUresult_t (*ModuleLoadData)(CUmodule_t*, const void*);
  CUresult_t (*ModuleGetFunction)(CUfunction_t*, CUmodule_t, const char*);
  CUresult_t (*ModuleUnload)(CUmodule_t);
  CUresult_t (*LaunchKernel)(CUfunction_t, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, unsigned, void*, void**, void**);
  CUresult_t (*FuncSetAttribute)(CUfunction_t, int, int);
  CUresult_t (*GetErrorString)(CUresult_t, const char**);
};

static TgNvrtc g_rtc = {};
static TgCuda g_cu = {};

static int tg_load_nvrtc() {
  if (g_rtc.h) return 0;
  const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
  void* h = nullptr;
  for (const char* n : names)
    if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) {
    tg_set_error("cannot dlopen libnvrtc: %s", dlerror());
    return 1;
  }
#define L(name)                                                        \
  *(void**)(&g_rtc.name) = dlsym(h, "nvrtc" #name);                    \
  if (!g_rtc.name) {                                                   \
    tg_set_error("libnvrtc lacks nvrtc" #name);                        \
    return 1;                                                          \
  }
  L(CreateProgram) L(CompileProgram) L(GetProgramLogSize) L(GetProgramLog) L(GetCUBINSize)
      L(GetCUBIN) L(DestroyProgram)
#undef L
  g_rtc.h = h;
  return 0;
}

static int tg_load_cuda() {
  if (g_cu.h) return 0;
  void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    tg_set_error("cannot dlopen libcuda.so.1 (no NVIDIA driver): %s", dlerror());
    return 1;
  }
#define L(field, sym)                                                  \
  *(void**)(&g_cu.field) = dlsym(h, sym);                              \
  if (!g_cu.field) {                                                   \
    tg_set_error("libcuda lacks " sym);                                \
    return 1;                                                          \
  }
  L(ModuleLoadData, "cuModuleLoadData") L(ModuleGetFunction, "cuModuleGetFunction")
      L(ModuleUnload, "cuModuleUnload") L(LaunchKernel, "cuLaunchKernel")
          L(FuncSetAttribute, "cuFuncSetAttribute") L(GetErrorString, "cuGetErrorString")
#undef L
  g_cu.h = h;
  return 0;
}

static const char* tg_cu_err(CUresult_t r) {
  const char* s = nullptr;
  if (g_cu.GetErrorString) g_cu.GetErrorString(r, &s);
  return s ? s : "unknown CUDA driver error";
}

// compile to an sm_100a cubin held in a malloc'd buffer
static int tg_jit_cubin(const char* src, std::vector<char>& cubin) {
  if (tg_load_nvrtc()) return 1;
  nvrtcProgram_t prog = nullptr;
  if (g_rtc.CreateProgram(&prog, src, "tigar_qp.cu", 0, nullptr, nullptr) != 0) {
    tg_set_error("nvrtcCreateProgram failed");
    return 1;
  }
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo",
                        "--fmad=true"};
  nvrtcResult_t rc = g_rtc.CompileProgram(prog, 4, opts);
  if (rc != 0) {
    size_t n = 0;
    g_rtc.GetProgramLogSize(prog, &n);
    std::string log(n + 1, '\0');
    if (n) g_rtc.GetProgramLog(prog, &log[0]);
    tg_set_error("nvrtc compile failed (%d): %.900s", rc, log.c_str());
    g_rtc.DestroyProgram(&prog);
    return 1;
  }
  size_t n = 0;
  if (g_rtc.GetCUBINSize(prog, &n) != 0 || n == 0) {
    tg_set_error("nvrtcGetCUBINSize failed");
    g_rtc.DestroyProgram(&prog);
    return 1;
  }
  cubin.resize(n);
  g_rtc.GetCUBIN(prog, cubin.data());
  g_rtc.DestroyProgram(&prog);
  return 0;
}

struct TgJitKernel {
  CUmodule_t mod;
  CUfunction_t fn;
};

// Compile only (works without a GPU): returns the cubin size.
extern "C" int tg_jit_check(const char* src, int64_t* cubin_bytes) {
  std::vector<char> cubin;
  if (tg_jit_cubin(src, cubin)) return 1;
  if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
  return 0;
}

extern "C" int tg_jit_compile(const char* src, const char* kernel_name, void** handle) {
  std::vector<char> cubin;
  if (tg_jit_cubin(src, cubin)) return 1;
  if (tg_load_cuda()) return 1;
  TG_CHECK(cudaFree(0));                       // make sure the primary context exists
  TgJitKernel* k = (TgJitKernel*)calloc(1, sizeof(TgJitKernel));
  CUresult_t r = g_cu.ModuleLoadData(&k->mod, cubin.data());
  if (r != 0) {
    tg_set_error("cuModuleLoadData: %s", tg_cu_err(r));
    free(k);
    return 1;
  }
  r = g_cu.ModuleGetFunction(&k->fn, k->mod, kernel_name);
  if (r != 0) {
    tg_set_error("cuModuleGetFunction(%s): %s", kernel_name, tg_cu_err(r));
    g_cu.ModuleUnload(k->mod);
    free(k);
    return 1;
  }
  *handle = k;
  return 0;
}

// One kernel parameter: a POD block passed by value (param, param_bytes).
extern "C" int tg_jit_launch(void* handle, int64_t grid, int32_t block, int32_t smem_bytes,
                             const void* param, int32_t param_bytes, void* stream) {
  TG_REQUIRE(handle != nullptr, "null JIT kernel");
  if (grid == 0) return 0;
  TG_REQUIRE(grid > 0 && grid < (int64_t)2147483647, "grid size");
  TgJitKernel* k = (TgJitKernel*)handle;
  (void)param_bytes;
  void* args[1] = {const_cast<void*>(param)};
  CUresult_t r = g_cu.LaunchKernel(k->fn, (unsigned)grid, 1, 1, (unsigned)block, 1, 1,
                                   (unsigned)smem_bytes, stream, args, nullptr);
  if (r != 0) {
    tg_set_error("cuLaunchKernel: %s", tg_cu_err(r));
    return 1;
  }
  tg_count_launch();
  return 0;
}

extern "C" int tg_jit_free(void* handle) {
  if (!handle) return 0;
  TgJitKernel* k = (TgJitKernel*)handle;
  if (g_cu.ModuleUnload) g_cu.ModuleUnload(k->mod);
  free(k);
  return 0;
}
