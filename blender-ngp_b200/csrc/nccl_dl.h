// NCCL, resolved at run time. The data-parallel path needs five NCCL entry points; they are looked up with dlopen so that
// (a) libngpb200.so loads on machines without NCCL as long as data parallelism is not used, and (b) inside a PyTorch process the copy
// of libnccl.so.2 that torch already loaded is the one used (dlopen by soname returns the loaded library) -- two NCCL runtimes in one
// process do not mix. Types follow nccl.h (2.27 / 2.28: ncclUniqueId is 128 bytes, ncclFloat = 7, ncclUint32 = 3, ncclSum = 0).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <stdexcept>
#include <string>

namespace ngpb {

struct NcclApi {
	typedef struct { char internal[128]; } UniqueId;
	typedef void* Comm;
	enum { Uint32 = 3, Float16 = 6, Float32 = 7, Bfloat16 = 9 }; // ncclDataType_t (nccl.h)
	enum { Sum = 0 };

	int (*GetUniqueId)(UniqueId*) = nullptr;
	int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
	int (*CommDestroy)(Comm) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
	int (*ReduceScatter)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
	int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;

	static NcclApi& get() {
		static NcclApi api = load();
		return api;
	}
	void check(int result, const char* what) const {
		if (result != 0) throw std::runtime_error(std::string(what) + " failed: " + (GetErrorString ? GetErrorString(result) : "NCCL error"));
	}

private:
	static NcclApi load() {
		void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!lib) throw std::runtime_error(std::string("data parallelism needs NCCL, and libnccl.so.2 could not be loaded: ") + dlerror());
		NcclApi a;
		auto sym = [&](const char* name) { void* p = dlsym(lib, name); if (!p) throw std::runtime_error(std::string("NCCL symbol missing: ") + name); return p; };
		a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
		a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
		a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
		a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
		a.ReduceScatter = reinterpret_cast<decltype(a.ReduceScatter)>(sym("ncclReduceScatter"));
		a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
		a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
		a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
		a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
		return a;
	}
};

} // namespace ngpb
