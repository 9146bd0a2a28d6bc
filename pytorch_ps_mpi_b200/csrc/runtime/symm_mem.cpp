#include "symm_mem.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <poll.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <chrono>
#include <stdexcept>

namespace psb {
namespace {

// Driver entry points are resolved through the runtime so the extension needs no -lcuda (the
// dev container has no driver; importing the module must still work there).
template <typename Fn>
Fn drv(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess)
    throw std::runtime_error(std::string("symm_mem: driver entry point not found: ") + name);
  return reinterpret_cast<Fn>(fn);
}

std::string cu_err(CUresult r) {
  static auto f = drv<CUresult (*)(CUresult, const char**)>("cuGetErrorString");
  const char* s = nullptr;
  f(r, &s);
  return s ? s : "unknown CUDA driver error";
}

#define CU_CHECK(call)                                                                             \
  do {                                                                                             \
    CUresult _r = (call);                                                                          \
    if (_r != CUDA_SUCCESS) throw std::runtime_error(std::string("symm_mem: " #call " failed: ") + cu_err(_r)); \
  } while (0)

struct Api {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*DeviceGet)(CUdevice*, int);
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                               unsigned long long);
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
};

const Api& api() {
  static Api a = [] {
    Api x{};
    x.MemCreate = drv<decltype(x.MemCreate)>("cuMemCreate");
    x.MemRelease = drv<decltype(x.MemRelease)>("cuMemRelease");
    x.MemGetAllocationGranularity = drv<decltype(x.MemGetAllocationGranularity)>("cuMemGetAllocationGranularity");
    x.MemExportToShareableHandle = drv<decltype(x.MemExportToShareableHandle)>("cuMemExportToShareableHandle");
    x.MemImportFromShareableHandle = drv<decltype(x.MemImportFromShareableHandle)>("cuMemImportFromShareableHandle");
    x.MemAddressReserve = drv<decltype(x.MemAddressReserve)>("cuMemAddressReserve");
    x.MemAddressFree = drv<decltype(x.MemAddressFree)>("cuMemAddressFree");
    x.MemMap = drv<decltype(x.MemMap)>("cuMemMap");
    x.MemUnmap = drv<decltype(x.MemUnmap)>("cuMemUnmap");
    x.MemSetAccess = drv<decltype(x.MemSetAccess)>("cuMemSetAccess");
    x.DeviceGet = drv<decltype(x.DeviceGet)>("cuDeviceGet");
    x.DeviceGetAttribute = drv<decltype(x.DeviceGetAttribute)>("cuDeviceGetAttribute");
    x.MulticastCreate = drv<decltype(x.MulticastCreate)>("cuMulticastCreate");
    x.MulticastAddDevice = drv<decltype(x.MulticastAddDevice)>("cuMulticastAddDevice");
    x.MulticastBindMem = drv<decltype(x.MulticastBindMem)>("cuMulticastBindMem");
    x.MulticastUnbind = drv<decltype(x.MulticastUnbind)>("cuMulticastUnbind");
    x.MulticastGetGranularity = drv<decltype(x.MulticastGetGranularity)>("cuMulticastGetGranularity");
    return x;
  }();
  return a;
}

size_t round_up(size_t n, size_t g) { return (n + g - 1) / g * g; }

sockaddr_un abstract_addr(const std::string& name, socklen_t* len) {
  sockaddr_un a{};
  a.sun_family = AF_UNIX;
  a.sun_path[0] = '\0';   // abstract namespace: nothing to unlink, vanishes with the process
  const size_t n = std::min(name.size(), sizeof(a.sun_path) - 2);
  memcpy(a.sun_path + 1, name.data(), n);
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
  return a;
}

void send_fd(int sock, int fd) {
  char payload = fd >= 0 ? 'F' : 'N';
  iovec io{&payload, 1};
  msghdr msg{};
  msg.msg_iov = &io;
  msg.msg_iovlen = 1;
  char ctl[CMSG_SPACE(sizeof(int))] = {0};
  if (fd >= 0) {
    msg.msg_control = ctl;
    msg.msg_controllen = sizeof(ctl);
    cmsghdr* c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
  }
  sendmsg(sock, &msg, 0);
}

int recv_fd(int sock) {
  char payload = 0;
  iovec io{&payload, 1};
  msghdr msg{};
  msg.msg_iov = &io;
  msg.msg_iovlen = 1;
  char ctl[CMSG_SPACE(sizeof(int))] = {0};
  msg.msg_control = ctl;
  msg.msg_controllen = sizeof(ctl);
  if (recvmsg(sock, &msg, 0) <= 0 || payload != 'F') return -1;
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return -1;
  int fd = -1;
  memcpy(&fd, CMSG_DATA(c), sizeof(int));
  return fd;
}

CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

uint64_t map_handle(CUmemGenericAllocationHandle h, size_t size, size_t gran, int device) {
  CUdeviceptr va = 0;
  CU_CHECK(api().MemAddressReserve(&va, size, gran, 0, 0));
  CU_CHECK(api().MemMap(va, size, 0, h, 0));
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CU_CHECK(api().MemSetAccess(va, size, &acc, 1));
  return (uint64_t)va;
}

}  // namespace

SymmBlock::SymmBlock(int rank, int world, int device, size_t bytes, const std::string& sock_prefix)
    : rank_(rank), world_(world), device_(device), requested_(bytes), prefix_(sock_prefix) {
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess)
    throw std::runtime_error("symm_mem: cannot initialise the CUDA context");
  CUmemAllocationProp prop = alloc_prop(device);
  size_t g = 0;
  CU_CHECK(api().MemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  gran_ = g;
  if (mc_supported()) {   // the multicast bind granularity can be larger than the allocation one
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)std::max(world, 1);
    mp.size = round_up(bytes, g);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (api().MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran_)
      gran_ = mg;
  }
  size_ = round_up(bytes ? bytes : 1, gran_);
  CUmemGenericAllocationHandle h = 0;
  CU_CHECK(api().MemCreate(&h, size_, &prop, 0));
  handle_ = h;
  ptrs_.assign(world, 0);
  peer_handles_.assign(world, 0);
  ptrs_[rank] = map_handle(h, size_, gran_, device);
  cudaMemset(reinterpret_cast<void*>(ptrs_[rank]), 0, size_);
  cudaDeviceSynchronize();
  if (world > 1) {
    int fd = -1;
    CU_CHECK(api().MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    local_fd_ = fd;
    listen_fd_ = socket(AF_UNIX, SOCK_STREAM, 0);
    socklen_t len;
    sockaddr_un a = abstract_addr(prefix_ + "-" + std::to_string(rank), &len);
    if (listen_fd_ < 0 || bind(listen_fd_, (sockaddr*)&a, len) != 0 || listen(listen_fd_, 64) != 0)
      throw std::runtime_error(std::string("symm_mem: cannot listen on the fd-exchange socket: ") + strerror(errno));
    server_ = std::thread([this] { serve(); });
  }
}

void SymmBlock::serve() {
  while (!stop_.load()) {
    pollfd p{listen_fd_, POLLIN, 0};
    if (poll(&p, 1, 50) <= 0) continue;
    int c = accept(listen_fd_, nullptr, nullptr);
    if (c < 0) continue;
    char what = 0;
    if (recv(c, &what, 1, 0) == 1) send_fd(c, what == 'm' ? local_fd_ : (what == 'c' ? mc_fd_ : -1));
    close(c);
  }
}

int SymmBlock::fetch_fd(int from_rank, char what) {
  socklen_t len;
  sockaddr_un a = abstract_addr(prefix_ + "-" + std::to_string(from_rank), &len);
  auto t0 = std::chrono::steady_clock::now();
  while (true) {
    int s = socket(AF_UNIX, SOCK_STREAM, 0);
    if (s >= 0 && connect(s, (sockaddr*)&a, len) == 0) {
      send(s, &what, 1, 0);
      int fd = recv_fd(s);
      close(s);
      if (fd >= 0) return fd;
    } else if (s >= 0) {
      close(s);
    }
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0)
      throw std::runtime_error("symm_mem: timed out fetching a handle from rank " + std::to_string(from_rank));
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
  }
}

void SymmBlock::map_peers() {
  cudaSetDevice(device_);
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    int fd = fetch_fd(r, 'm');
    CUmemGenericAllocationHandle h = 0;
    CUresult res = api().MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    if (res != CUDA_SUCCESS) throw std::runtime_error("symm_mem: import of rank " + std::to_string(r) + " failed: " + cu_err(res));
    peer_handles_[r] = h;
    ptrs_[r] = map_handle(h, size_, gran_, device_);
  }
}

bool SymmBlock::mc_supported() const {
  try {
    CUdevice d;
    if (api().DeviceGet(&d, device_) != CUDA_SUCCESS) return false;
    int v = 0;
    if (api().DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d) != CUDA_SUCCESS) return false;
    return v != 0;
  } catch (...) {
    return false;
  }
}

bool SymmBlock::mc_create() {
  try {
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)world_;
    mp.size = size_;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle h = 0;
    CUresult r = api().MulticastCreate(&h, &mp);
    if (r != CUDA_SUCCESS) {
      err_ = "cuMulticastCreate: " + cu_err(r);
      return false;
    }
    mc_handle_ = h;
    mc_created_ = true;
    int fd = -1;
    r = api().MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) {
      err_ = "export multicast handle: " + cu_err(r);
      return false;
    }
    mc_fd_ = fd;
    return true;
  } catch (const std::exception& e) {
    err_ = e.what();
    return false;
  }
}

bool SymmBlock::mc_import() {
  try {
    int fd = fetch_fd(0, 'c');
    CUmemGenericAllocationHandle h = 0;
    CUresult r = api().MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    if (r != CUDA_SUCCESS) {
      err_ = "import multicast handle: " + cu_err(r);
      return false;
    }
    mc_handle_ = h;
    mc_created_ = true;
    return true;
  } catch (const std::exception& e) {
    err_ = e.what();
    return false;
  }
}

bool SymmBlock::mc_add_device() {
  CUdevice d;
  if (!mc_created_ || api().DeviceGet(&d, device_) != CUDA_SUCCESS) return false;
  CUresult r = api().MulticastAddDevice(mc_handle_, d);
  if (r != CUDA_SUCCESS) {
    err_ = "cuMulticastAddDevice: " + cu_err(r);
    return false;
  }
  return true;
}

bool SymmBlock::mc_bind_and_map() {
  if (!mc_created_) return false;
  CUresult r = api().MulticastBindMem(mc_handle_, 0, handle_, 0, size_, 0);
  if (r != CUDA_SUCCESS) {
    err_ = "cuMulticastBindMem: " + cu_err(r);
    return false;
  }
  mc_bound_ = true;
  try {
    mc_ptr_ = map_handle(mc_handle_, size_, gran_, device_);
  } catch (const std::exception& e) {
    err_ = e.what();
    mc_ptr_ = 0;
    return false;
  }
  return true;
}

void SymmBlock::stop_server() {
  stop_.store(true);
  if (server_.joinable()) server_.join();
  if (listen_fd_ >= 0) close(listen_fd_), listen_fd_ = -1;
  if (local_fd_ >= 0) close(local_fd_), local_fd_ = -1;
  if (mc_fd_ >= 0) close(mc_fd_), mc_fd_ = -1;
}

SymmBlock::~SymmBlock() {
  stop_server();
  // Best effort; errors during interpreter teardown are ignored.
  try {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    if (mc_ptr_) {
      api().MemUnmap((CUdeviceptr)mc_ptr_, size_);
      api().MemAddressFree((CUdeviceptr)mc_ptr_, size_);
    }
    if (mc_bound_) {
      CUdevice d;
      if (api().DeviceGet(&d, device_) == CUDA_SUCCESS) api().MulticastUnbind(mc_handle_, d, 0, size_);
    }
    if (mc_created_) api().MemRelease(mc_handle_);
    for (int r = 0; r < world_; ++r) {
      if (ptrs_[r]) {
        api().MemUnmap((CUdeviceptr)ptrs_[r], size_);
        api().MemAddressFree((CUdeviceptr)ptrs_[r], size_);
      }
      if (r != rank_ && peer_handles_[r]) api().MemRelease(peer_handles_[r]);
    }
    if (handle_) api().MemRelease(handle_);
  } catch (...) {
  }
}

}  // namespace psb
