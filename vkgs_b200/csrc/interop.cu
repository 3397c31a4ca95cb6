// External-memory / external-semaphore side of the C ABI (the present path of a desktop viewer).
//
// The reference's own (unbuilt) interop code creates a VkImage whose memory is exported as an opaque file descriptor,
// imports it into CUDA and maps a buffer onto it (src/vkgs/engine/interop/cuda_image.cu:77-132), and does the same for a
// VkSemaphore (interop/cuda_semaphore.cu:56-86; the device extensions are enabled at vulkan/context.cc:204-216,240-241).
// Here the CUDA half of that contract is an entry point: hand over the fd, get a device pointer, pass it to
// vkgsb_draw(dst, dst_is_device = 1) - the blend kernel then writes the frame straight into the Vulkan image's memory -
// and signal the imported semaphore on the frame's stream.
//
// No Vulkan loader exists in this image or on the GPU box (profiles/r02_host_vulkan_probe.txt), so the opaque-fd path
// cannot be exercised against a real VkDeviceMemory here.  The same file-descriptor contract is tested with the one
// exporter that does exist without Vulkan: a CUDA virtual-memory allocation exported as a POSIX fd
// (cuMemExportToShareableHandle).  vkgsb_external_alloc() creates such an allocation (the stand-in for the VkImage's
// memory: another process, or another API in this process, owns it and passes the fd), vkgsb_external_import() with
// VKGSB_EXTERNAL_CUDA_POSIX_FD maps it (cuMemImportFromShareableHandle + reserve / map / set-access).  The driver entry
// points are taken from the runtime (cudaGetDriverEntryPoint): the library carries no link-time dependency on libcuda.
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstring>
#include <string>

#include "../../include/vkgsb.h"

extern "C" void vkgsb_set_last_error(const char* msg);  // renderer.cu

namespace {

int fail(int code, const std::string& msg) {
  vkgsb_set_last_error(msg.c_str());
  return code;
}

#define RT_TRY(expr)                                                                               \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(VKGSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// the driver API entry points this file needs, resolved through the runtime
struct Driver {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};

template <class F>
bool entry(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
  *fn = reinterpret_cast<F>(p);
  return true;
}

const Driver& driver() {
  static Driver d = [] {
    Driver x;
    x.ok = entry("cuMemCreate", &x.MemCreate) && entry("cuMemRelease", &x.MemRelease) &&
           entry("cuMemGetAllocationGranularity", &x.MemGetAllocationGranularity) &&
           entry("cuMemExportToShareableHandle", &x.MemExportToShareableHandle) &&
           entry("cuMemImportFromShareableHandle", &x.MemImportFromShareableHandle) &&
           entry("cuMemAddressReserve", &x.MemAddressReserve) && entry("cuMemAddressFree", &x.MemAddressFree) &&
           entry("cuMemMap", &x.MemMap) && entry("cuMemUnmap", &x.MemUnmap) && entry("cuMemSetAccess", &x.MemSetAccess) &&
           entry("cuGetErrorString", &x.GetErrorString);
    return x;
  }();
  return d;
}

std::string cu_text(CUresult r) {
  const char* s = nullptr;
  if (driver().GetErrorString) driver().GetErrorString(r, &s);
  return s ? s : "CUDA driver error " + std::to_string(static_cast<int>(r));
}

#define DRV_TRY(expr)                                                                          \
  do {                                                                                         \
    CUresult r_ = (expr);                                                                      \
    if (r_ != CUDA_SUCCESS) return fail(VKGSB_ERR_CUDA, std::string(#expr) + ": " + cu_text(r_)); \
  } while (0)

CUmemAllocationProp vmm_prop(int device) {
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

}  // namespace

struct vkgsb_external {
  int device = 0;
  int kind = 0;                         // vkgsb_external_handle_type
  size_t bytes = 0, mapped = 0;         // requested / rounded up to the allocation granularity
  cudaExternalMemory_t ext = nullptr;   // VKGSB_EXTERNAL_OPAQUE_FD
  void* ptr = nullptr;
  CUmemGenericAllocationHandle handle = 0;  // VKGSB_EXTERNAL_CUDA_POSIX_FD
  bool have_handle = false;
};

namespace {

int vmm_map(vkgsb_external* x) {
  const Driver& d = driver();
  CUdeviceptr va = 0;
  DRV_TRY(d.MemAddressReserve(&va, x->mapped, 0, 0, 0));
  CUresult r = d.MemMap(va, x->mapped, 0, x->handle, 0);
  if (r == CUDA_SUCCESS) {
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = x->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = d.MemSetAccess(va, x->mapped, &acc, 1);
    if (r != CUDA_SUCCESS) d.MemUnmap(va, x->mapped);
  }
  if (r != CUDA_SUCCESS) {
    d.MemAddressFree(va, x->mapped);
    return fail(VKGSB_ERR_CUDA, "mapping the imported allocation: " + cu_text(r));
  }
  x->ptr = reinterpret_cast<void*>(va);
  return VKGSB_OK;
}

}  // namespace

extern "C" {

int vkgsb_external_alloc(int device, size_t bytes, vkgsb_external** out, int* fd, void** d_ptr) {
  if (!out || !fd || !d_ptr || bytes == 0) return fail(VKGSB_ERR_INVALID, "null argument");
  RT_TRY(cudaSetDevice(device));
  RT_TRY(cudaFree(nullptr));  // the context exists before the driver API is used
  const Driver& d = driver();
  if (!d.ok) return fail(VKGSB_ERR_CUDA, "the CUDA driver's virtual-memory entry points are not available");
  const CUmemAllocationProp prop = vmm_prop(device);
  size_t gran = 0;
  DRV_TRY(d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  auto* x = new vkgsb_external();
  x->device = device;
  x->kind = VKGSB_EXTERNAL_CUDA_POSIX_FD;
  x->bytes = bytes;
  x->mapped = (bytes + gran - 1) / gran * gran;
  CUresult r = d.MemCreate(&x->handle, x->mapped, &prop, 0);
  if (r != CUDA_SUCCESS) {
    delete x;
    return fail(VKGSB_ERR_CUDA, "cuMemCreate (exportable): " + cu_text(r));
  }
  x->have_handle = true;
  int f = -1;
  r = d.MemExportToShareableHandle(&f, x->handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    vkgsb_external_release(x);
    return fail(VKGSB_ERR_CUDA, "cuMemExportToShareableHandle: " + cu_text(r));
  }
  if (int e = vmm_map(x)) {
    close(f);
    vkgsb_external_release(x);
    return e;
  }
  *out = x;
  *fd = f;
  *d_ptr = x->ptr;
  return VKGSB_OK;
}

int vkgsb_external_import(int device, int fd, size_t bytes, int handle_type, vkgsb_external** out, void** d_ptr) {
  if (!out || !d_ptr || bytes == 0) return fail(VKGSB_ERR_INVALID, "null argument");
  if (fd < 0) return fail(VKGSB_ERR_INVALID, "bad file descriptor");
  RT_TRY(cudaSetDevice(device));
  RT_TRY(cudaFree(nullptr));
  auto* x = new vkgsb_external();
  x->device = device;
  x->kind = handle_type;
  x->bytes = bytes;
  if (handle_type == VKGSB_EXTERNAL_OPAQUE_FD) {
    // cuda_image.cu:33-66 of the reference: import the VkDeviceMemory's fd, map a buffer onto it.  On success the fd
    // belongs to CUDA (it must not be closed by the caller).
    cudaExternalMemoryHandleDesc hd{};
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    hd.size = bytes;
    cudaError_t e = cudaImportExternalMemory(&x->ext, &hd);
    if (e != cudaSuccess) {
      cudaGetLastError();
      delete x;
      return fail(VKGSB_ERR_CUDA, std::string("cudaImportExternalMemory (opaque fd): ") + cudaGetErrorString(e));
    }
    cudaExternalMemoryBufferDesc bd{};
    bd.offset = 0;
    bd.size = bytes;
    e = cudaExternalMemoryGetMappedBuffer(&x->ptr, x->ext, &bd);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cudaDestroyExternalMemory(x->ext);
      delete x;
      return fail(VKGSB_ERR_CUDA, std::string("cudaExternalMemoryGetMappedBuffer: ") + cudaGetErrorString(e));
    }
  } else if (handle_type == VKGSB_EXTERNAL_CUDA_POSIX_FD) {
    const Driver& d = driver();
    if (!d.ok) {
      delete x;
      return fail(VKGSB_ERR_CUDA, "the CUDA driver's virtual-memory entry points are not available");
    }
    const CUmemAllocationProp prop = vmm_prop(device);
    size_t gran = 0;
    CUresult r = d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    if (r == CUDA_SUCCESS) {
      x->mapped = (bytes + gran - 1) / gran * gran;
      r = d.MemImportFromShareableHandle(&x->handle, reinterpret_cast<void*>(static_cast<intptr_t>(fd)),
                                         CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    }
    if (r != CUDA_SUCCESS) {
      delete x;
      return fail(VKGSB_ERR_CUDA, "cuMemImportFromShareableHandle: " + cu_text(r));
    }
    x->have_handle = true;
    if (int e = vmm_map(x)) {
      vkgsb_external_release(x);
      return e;
    }
  } else {
    delete x;
    return fail(VKGSB_ERR_INVALID, "unknown external handle type");
  }
  *out = x;
  *d_ptr = x->ptr;
  return VKGSB_OK;
}

int vkgsb_external_release(vkgsb_external* x) {
  if (!x) return VKGSB_OK;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  if (x->kind == VKGSB_EXTERNAL_OPAQUE_FD) {
    if (x->ptr) cudaFree(x->ptr);  // "must eventually be freed using cudaFree" (cuda_image.cu:63)
    if (x->ext) cudaDestroyExternalMemory(x->ext);
  } else {
    const Driver& d = driver();
    if (x->ptr) {
      d.MemUnmap(reinterpret_cast<CUdeviceptr>(x->ptr), x->mapped);
      d.MemAddressFree(reinterpret_cast<CUdeviceptr>(x->ptr), x->mapped);
    }
    if (x->have_handle) d.MemRelease(x->handle);
  }
  delete x;
  return VKGSB_OK;
}

// ---- semaphores (cuda_semaphore.cu:29-86): an opaque fd exported from a VkSemaphore; signalled behind a frame so that the
//      Vulkan queue that presents the image waits for it, waited on before a frame overwrites an image still being read.
int vkgsb_external_semaphore_import(int device, int fd, void** sem) {
  if (!sem) return fail(VKGSB_ERR_INVALID, "null argument");
  if (fd < 0) return fail(VKGSB_ERR_INVALID, "bad file descriptor");
  RT_TRY(cudaSetDevice(device));
  cudaExternalSemaphoreHandleDesc hd{};
  hd.type = cudaExternalSemaphoreHandleTypeOpaqueFd;
  hd.handle.fd = fd;
  cudaExternalSemaphore_t s = nullptr;
  cudaError_t e = cudaImportExternalSemaphore(&s, &hd);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(VKGSB_ERR_CUDA, std::string("cudaImportExternalSemaphore (opaque fd): ") + cudaGetErrorString(e));
  }
  *sem = s;
  return VKGSB_OK;
}

int vkgsb_external_semaphore_signal(void* sem, void* stream) {
  if (!sem) return fail(VKGSB_ERR_INVALID, "null semaphore");
  cudaExternalSemaphore_t s = static_cast<cudaExternalSemaphore_t>(sem);
  cudaExternalSemaphoreSignalParams p{};
  RT_TRY(cudaSignalExternalSemaphoresAsync(&s, &p, 1, static_cast<cudaStream_t>(stream)));
  return VKGSB_OK;
}

int vkgsb_external_semaphore_wait(void* sem, void* stream) {
  if (!sem) return fail(VKGSB_ERR_INVALID, "null semaphore");
  cudaExternalSemaphore_t s = static_cast<cudaExternalSemaphore_t>(sem);
  cudaExternalSemaphoreWaitParams p{};
  RT_TRY(cudaWaitExternalSemaphoresAsync(&s, &p, 1, static_cast<cudaStream_t>(stream)));
  return VKGSB_OK;
}

int vkgsb_external_semaphore_release(void* sem) {
  if (!sem) return VKGSB_OK;
  RT_TRY(cudaDestroyExternalSemaphore(static_cast<cudaExternalSemaphore_t>(sem)));
  return VKGSB_OK;
}

}  // extern "C"
