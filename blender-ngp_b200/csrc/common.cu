#include "common.cuh"
#include "../../include/ngpb.h"

#include <mutex>

namespace ngpb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
} // namespace ngpb

extern "C" const char* ngpb_last_error(void) { return ngpb::g_last_error.c_str(); }
extern "C" int ngpb_version(void) { return 100; }

extern "C" int ngpb_check_device(int device) {
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess) { ngpb::set_last_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); return (int)e; }
	if (device < 0 || device >= count) { ngpb::set_last_error("no such CUDA device"); return NGPB_ERR_INVALID_ARGUMENT; }
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess) { ngpb::set_last_error(cudaGetErrorString(e)); return (int)e; }
	if (prop.major != 10) {
		ngpb::set_last_error(std::string("libngpb200 is built for sm_100a only; device is ") + prop.name + " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
		return NGPB_ERR_RUNTIME;
	}
	return 0;
}
