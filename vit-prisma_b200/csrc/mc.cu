// mc.cu -- NVSwitch multicast (NVLS) memory for the data-parallel SAE step: one physical allocation per rank, all bound to one
// multicast object, so that
//     multimem.ld_reduce  on the multicast address returns the SUM over every rank's copy (the reduction happens in the switch:
//                         a rank pulls its 1/N gradient slice ONCE instead of once per peer), and
//     multimem.st         writes every rank's copy with one store (all-gather of updated parameters: 1/N slice out, once).
// Link bytes per GPU per step fall from 2 (N-1)/N x 8 d F  to  2 x 8 d F / N.
//
// Driver-API plumbing (no NCCL): cuMulticastCreate on rank 0 -> POSIX file descriptor -> the other ranks import it (the fd travels
// over a Unix-domain socket, SCM_RIGHTS, in vit_prisma/b200/p2p.py) -> every rank cuMulticastAddDevice -> [barrier] -> every rank
// cuMemCreate (shareable) + cuMulticastBindMem at offset 0 -> [barrier] -> unicast and multicast mappings.
// Everything here fails softly (PB_EUNSUPPORTED + message): the caller falls back to the peer load / store path of p2p.cu.
#include "common.cuh"
#include <cuda.h>

namespace {

struct McApi {
  bool ok = false;
  CUresult (*DeviceGet)(CUdevice*, int);
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*GetErrorString)(CUresult, const char**);
};

template <typename F>
bool mc_load(F& fn, const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
  fn = reinterpret_cast<F>(p);
  return true;
}

McApi& api() {
  static McApi a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    a.ok = mc_load(a.DeviceGet, "cuDeviceGet") && mc_load(a.DeviceGetAttribute, "cuDeviceGetAttribute") &&
           mc_load(a.MulticastGetGranularity, "cuMulticastGetGranularity") && mc_load(a.MulticastCreate, "cuMulticastCreate") &&
           mc_load(a.MulticastAddDevice, "cuMulticastAddDevice") && mc_load(a.MulticastBindMem, "cuMulticastBindMem") &&
           mc_load(a.MemCreate, "cuMemCreate") && mc_load(a.MemRelease, "cuMemRelease") &&
           mc_load(a.MemExportToShareableHandle, "cuMemExportToShareableHandle") &&
           mc_load(a.MemImportFromShareableHandle, "cuMemImportFromShareableHandle") && mc_load(a.MemAddressReserve, "cuMemAddressReserve") &&
           mc_load(a.MemAddressFree, "cuMemAddressFree") && mc_load(a.MemMap, "cuMemMap") && mc_load(a.MemUnmap, "cuMemUnmap") &&
           mc_load(a.MemSetAccess, "cuMemSetAccess") && mc_load(a.MemGetAllocationGranularity, "cuMemGetAllocationGranularity") &&
           mc_load(a.GetErrorString, "cuGetErrorString");
  }
  return a;
}

int mc_fail(const char* what, CUresult rc) {
  const char* msg = nullptr;
  if (api().ok) api().GetErrorString(rc, &msg);
  pb_set_error("multicast: %s failed (%d: %s)", what, (int)rc, msg ? msg : "?");
  return PB_EUNSUPPORTED;
}
#define MC_DRV(call, what)                       \
  do {                                           \
    CUresult rc__ = (call);                      \
    if (rc__ != CUDA_SUCCESS) return mc_fail(what, rc__); \
  } while (0)

int current_device(CUdevice* dev) {
  int ord = 0;
  PB_CUDA(cudaGetDevice(&ord));
  PB_CUDA(cudaFree(0));                           // make sure the primary context exists
  MC_DRV(api().DeviceGet(dev, ord), "cuDeviceGet");
  return PB_OK;
}

CUmulticastObjectProp mc_prop(int world, size_t bytes) {
  CUmulticastObjectProp p;
  memset(&p, 0, sizeof(p));
  p.numDevices = (unsigned)world;
  p.size = bytes;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

}  // namespace

// *supported = 1 when the current device can join a multicast team (NVSwitch fabric with multicast enabled)
extern "C" int pb_mc_supported(int32_t* supported) {
  PB_CHECK_ARG(supported, "pb_mc_supported: null argument");
  *supported = 0;
  if (!api().ok) return PB_OK;
  CUdevice dev;
  if (current_device(&dev) != PB_OK) return PB_OK;
  int v = 0;
  if (api().DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS) *supported = v ? 1 : 0;
  return PB_OK;
}

// size of the multicast object / of every rank's allocation for a request of `bytes` (rounded up to the recommended granularity)
extern "C" int pb_mc_round_size(int32_t world, int64_t bytes, int64_t* rounded) {
  PB_CHECK_ARG(world >= 1 && bytes > 0 && rounded, "pb_mc_round_size: bad arguments");
  if (!api().ok) { pb_set_error("multicast: driver entry points unavailable"); return PB_EUNSUPPORTED; }
  CUmulticastObjectProp p = mc_prop(world, (size_t)bytes);
  size_t gran = 0;
  MC_DRV(api().MulticastGetGranularity(&gran, &p, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
  CUdevice dev;
  PB_TRY(current_device(&dev));
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = (int)dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t g2 = 0;
  MC_DRV(api().MemGetAllocationGranularity(&g2, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  if (g2 > gran) gran = g2;
  *rounded = (int64_t)(((size_t)bytes + gran - 1) / gran * gran);
  return PB_OK;
}

// rank 0: create the multicast object (size from pb_mc_round_size) and export it as a POSIX file descriptor
extern "C" int pb_mc_create(int32_t world, int64_t bytes, uint64_t* mc_handle, int32_t* fd) {
  PB_CHECK_ARG(world >= 1 && bytes > 0 && mc_handle && fd, "pb_mc_create: bad arguments");
  if (!api().ok) { pb_set_error("multicast: driver entry points unavailable"); return PB_EUNSUPPORTED; }
  CUmulticastObjectProp p = mc_prop(world, (size_t)bytes);
  CUmemGenericAllocationHandle h;
  MC_DRV(api().MulticastCreate(&h, &p), "cuMulticastCreate");
  int f = -1;
  MC_DRV(api().MemExportToShareableHandle(&f, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
  *mc_handle = (uint64_t)h;
  *fd = f;
  return PB_OK;
}

// other ranks: import the multicast object from the file descriptor received from rank 0
extern "C" int pb_mc_import(int32_t fd, uint64_t* mc_handle) {
  PB_CHECK_ARG(fd >= 0 && mc_handle, "pb_mc_import: bad arguments");
  if (!api().ok) { pb_set_error("multicast: driver entry points unavailable"); return PB_EUNSUPPORTED; }
  CUmemGenericAllocationHandle h;
  MC_DRV(api().MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle");
  *mc_handle = (uint64_t)h;
  return PB_OK;
}

// every rank, before anybody binds: join the team with the current device
extern "C" int pb_mc_add_device(uint64_t mc_handle) {
  if (!api().ok) { pb_set_error("multicast: driver entry points unavailable"); return PB_EUNSUPPORTED; }
  CUdevice dev;
  PB_TRY(current_device(&dev));
  MC_DRV(api().MulticastAddDevice((CUmemGenericAllocationHandle)mc_handle, dev), "cuMulticastAddDevice");
  return PB_OK;
}

// every rank, after ALL ranks added their device: allocate `bytes` of device memory (zeroed), bind it at offset 0 of the
// multicast object and map it twice: *uc_ptr = this rank's own copy (ordinary loads / stores), *mc_ptr = the multicast view
// (multimem.* instructions only; valid once every rank has bound).
extern "C" int pb_mc_bind_alloc(uint64_t mc_handle, int64_t bytes, void** uc_ptr, void** mc_ptr, uint64_t* mem_handle) {
  PB_CHECK_ARG(bytes > 0 && uc_ptr && mc_ptr && mem_handle, "pb_mc_bind_alloc: bad arguments");
  if (!api().ok) { pb_set_error("multicast: driver entry points unavailable"); return PB_EUNSUPPORTED; }
  CUdevice dev;
  PB_TRY(current_device(&dev));
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = (int)dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle mem;
  MC_DRV(api().MemCreate(&mem, (size_t)bytes, &ap, 0), "cuMemCreate");
  MC_DRV(api().MulticastBindMem((CUmemGenericAllocationHandle)mc_handle, 0, mem, 0, (size_t)bytes, 0), "cuMulticastBindMem");
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = (int)dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CUdeviceptr uva = 0, mva = 0;
  MC_DRV(api().MemAddressReserve(&uva, (size_t)bytes, 0, 0, 0), "cuMemAddressReserve(unicast)");
  MC_DRV(api().MemMap(uva, (size_t)bytes, 0, mem, 0), "cuMemMap(unicast)");
  MC_DRV(api().MemSetAccess(uva, (size_t)bytes, &acc, 1), "cuMemSetAccess(unicast)");
  MC_DRV(api().MemAddressReserve(&mva, (size_t)bytes, 0, 0, 0), "cuMemAddressReserve(multicast)");
  MC_DRV(api().MemMap(mva, (size_t)bytes, 0, (CUmemGenericAllocationHandle)mc_handle, 0), "cuMemMap(multicast)");
  MC_DRV(api().MemSetAccess(mva, (size_t)bytes, &acc, 1), "cuMemSetAccess(multicast)");
  PB_CUDA(cudaMemset((void*)uva, 0, (size_t)bytes));
  PB_CUDA(cudaDeviceSynchronize());
  *uc_ptr = (void*)uva;
  *mc_ptr = (void*)mva;
  *mem_handle = (uint64_t)mem;
  return PB_OK;
}
