// abi.cu -- version / error-string entry points of libpcnbr.
#include "common.cuh"

extern "C" int pcnbr_abi_version(void) { return PCNBR_ABI_VERSION; }

extern "C" const char* pcnbr_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case PCNBR_E_BADARG: return "pcnbr: bad argument (null pointer, non-positive size or K > N)";
        case PCNBR_E_TOOLARGE: return "pcnbr: size outside the compiled limits (K <= 128, F <= 256, pool K <= 255, interp k <= 8)";
        case PCNBR_E_WORKSPACE: return "pcnbr: workspace missing or too small";
        default: return (code > 0) ? cudaGetErrorString((cudaError_t)code) : "pcnbr: unknown error";
    }
}
