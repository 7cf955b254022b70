// Library-level entry points: ABI version, error strings, device capability probe.
#include "common.cuh"

extern "C" int dkt_abi_version(void) { return DKT_ABI_VERSION; }

extern "C" int dkt_split_format(void) { return DKT_SPLIT_FP16 ? DKT_FMT_FP16 : DKT_FMT_BF16; }

extern "C" const char* dkt_error_string(int code) {
    switch (code) {
        case 0:                 return "ok";
        case DKT_E_INVALID:     return "invalid argument (null pointer or bad dimension)";
        case DKT_E_UNSUPPORTED: return "shape or option outside what the sm_100a kernels support";
        case DKT_E_ALIGNMENT:   return "pointer or channel count not aligned as the kernel requires";
        case DKT_E_DRIVER:      return "CUDA driver entry point (cuTensorMapEncodeTiled) unavailable or failed";
        default:
            if (code > 0) return cudaGetErrorString((cudaError_t)code);
            return "unknown dkt error";
    }
}

extern "C" int dkt_device_supported(int device) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { cudaGetLastError(); return -(int)e - 100; }
    return prop.major == 10 ? 1 : 0;
}
