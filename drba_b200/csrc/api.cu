// Library-level entry points of libdrba_b200.so.
#include "common.cuh"

extern "C" {

int drba_version(void) { return 100; }  // 0.1.0

const char* drba_error_string(int code)
{
    switch (code) {
        case DRBA_OK: return "ok";
        case DRBA_E_ARG: return "invalid argument (NULL pointer, negative dimension or bad enum)";
        case DRBA_E_WORKSPACE: return "workspace missing or too small";
        case DRBA_E_UNSUPPORTED: return "unsupported configuration";
        case DRBA_E_ALIGN: return "pointer is not 16-byte aligned";
        case DRBA_E_BARRIER: return "a persistent conv program timed out at its grid barrier (CTAs not co-resident): results are invalid";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown drba error";
}

int drba_workspace_clear(void* ws, size_t bytes, void* stream)
{
    if (bytes == 0) return DRBA_OK;
    if (!ws) return DRBA_E_ARG;
    const cudaError_t e = cudaMemsetAsync(ws, 0, bytes, drba::as_stream(stream));
    return e == cudaSuccess ? DRBA_OK : (int)e;
}

}  // extern "C"
