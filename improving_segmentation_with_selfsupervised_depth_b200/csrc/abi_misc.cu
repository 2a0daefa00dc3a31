// Library-level ABI entry points (version, error strings, launch counter).
#include "common.cuh"
namespace segsde { std::atomic<int64_t> g_launches{0}; }
extern "C" const char* segsde_version(void) { return "segsde_b200 0.1 (sm_100a)"; }
extern "C" int64_t segsde_launch_count(void) { return segsde::g_launches.load(); }
extern "C" const char* segsde_error_string(int code) {
  switch (code) {
    case SEGSDE_OK: return "ok";
    case SEGSDE_E_ARG: return "segsde: invalid argument (shape / null pointer / flag combination)";
    case SEGSDE_E_ALIGN: return "segsde: pointer or stride alignment requirement violated";
    case SEGSDE_E_UNSUPPORTED: return "segsde: configuration not supported by this kernel family";
    case SEGSDE_E_WORKSPACE: return "segsde: workspace too small";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "segsde: unknown error";
  }
}
