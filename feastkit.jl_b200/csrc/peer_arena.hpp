// Shared arena of a row-sharded run: one physical allocation per rank (CUDA virtual-memory-management API), exported as a POSIX file
// descriptor, passed to the peer processes over Unix-domain sockets (SCM_RIGHTS) and mapped by EVERY process into one contiguous
// virtual range -- rank p's arena sits at base + p * stride.  A halo row in a peer's HBM is then a fixed SIGNED 32-bit distance (in
// 16-byte units) from the corresponding local row, so the gather kernels address local and remote rows with the same instruction
// (the loads of a remote row travel over NVLink).  Driver entry points are resolved through cudaGetDriverEntryPoint: the library keeps
// no link-time dependency on libcuda and still loads on a machine without a GPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace feastcuda {

struct VmmApi {
  CUresult (*GetGran)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*Export)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*Import)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*Reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
  bool ok = false;
};

inline VmmApi& vmm_api() {
  static VmmApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  auto get = [](const char* name, void** fn) {
    cudaDriverEntryPointQueryResult st;
    return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess && *fn != nullptr;
  };
  api.ok = get("cuMemGetAllocationGranularity", (void**)&api.GetGran) && get("cuMemCreate", (void**)&api.Create) &&
           get("cuMemExportToShareableHandle", (void**)&api.Export) && get("cuMemImportFromShareableHandle", (void**)&api.Import) &&
           get("cuMemAddressReserve", (void**)&api.Reserve) && get("cuMemMap", (void**)&api.Map) &&
           get("cuMemSetAccess", (void**)&api.SetAccess) && get("cuMemUnmap", (void**)&api.Unmap) &&
           get("cuMemAddressFree", (void**)&api.AddressFree) && get("cuMemRelease", (void**)&api.Release);
  return api;
}

struct PeerArena {
  CUdeviceptr base = 0;          // start of the contiguous range (nranks * stride bytes)
  size_t stride = 0;             // bytes per rank (multiple of the allocation granularity)
  int nranks = 0, rank = 0;
  std::vector<CUmemGenericAllocationHandle> handles;   // [nranks], handles[rank] = the local allocation
  std::vector<char> mapped;
  bool active = false;
};

// ---- file descriptors between the ranks of one node: abstract-namespace datagram sockets, one per rank ----------------------------
inline void fd_addr(sockaddr_un& a, socklen_t& len, unsigned long long token, int rank) {
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  char name[64];
  const int k = snprintf(name, sizeof(name), "feastcuda-%016llx-%d", token, rank);
  a.sun_path[0] = '\0';                                   // abstract namespace: no file system entry to clean up
  memcpy(a.sun_path + 1, name, (size_t)k);
  len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + k);
}
inline int fd_socket_open(unsigned long long token, int rank) {
  const int s = socket(AF_UNIX, SOCK_DGRAM, 0);
  if (s < 0) return -1;
  sockaddr_un a;
  socklen_t len;
  fd_addr(a, len, token, rank);
  if (bind(s, (sockaddr*)&a, len) != 0) { close(s); return -1; }
  return s;
}
inline bool fd_send(int s, unsigned long long token, int to_rank, int my_rank, int fd) {
  sockaddr_un a;
  socklen_t len;
  fd_addr(a, len, token, to_rank);
  int payload = my_rank;
  iovec iov{&payload, sizeof(payload)};
  char ctl[CMSG_SPACE(sizeof(int))];
  memset(ctl, 0, sizeof(ctl));
  msghdr msg{};
  msg.msg_name = &a;
  msg.msg_namelen = len;
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctl;
  msg.msg_controllen = sizeof(ctl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  return sendmsg(s, &msg, 0) == (ssize_t)sizeof(payload);
}
inline bool fd_recv(int s, int* from_rank, int* fd) {
  int payload = -1;
  iovec iov{&payload, sizeof(payload)};
  char ctl[CMSG_SPACE(sizeof(int))];
  memset(ctl, 0, sizeof(ctl));
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctl;
  msg.msg_controllen = sizeof(ctl);
  if (recvmsg(s, &msg, 0) != (ssize_t)sizeof(payload)) return false;
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return false;
  memcpy(fd, CMSG_DATA(c), sizeof(int));
  *from_rank = payload;
  return true;
}

inline void peer_arena_destroy(PeerArena& pa) {
  VmmApi& api = vmm_api();
  if (!pa.active) return;
  for (int p = 0; p < pa.nranks; ++p) {
    if (pa.mapped[p]) api.Unmap(pa.base + (size_t)p * pa.stride, pa.stride);
    if (pa.handles[p]) api.Release(pa.handles[p]);
  }
  if (pa.base) api.AddressFree(pa.base, pa.stride * (size_t)pa.nranks);
  pa = PeerArena();
}

// barrier(): a collective the caller provides (every rank has bound its socket / received every descriptor).
// Returns an empty string on success, else what failed (the caller reports it).
template <typename Barrier>
inline std::string peer_arena_create(PeerArena& pa, int device, int nranks, int rank, unsigned long long token, size_t bytes, Barrier barrier) {
  VmmApi& api = vmm_api();
  if (!api.ok) return "CUDA virtual-memory-management entry points unavailable";
  pa = PeerArena();
  pa.nranks = nranks;
  pa.rank = rank;
  pa.handles.assign(nranks, 0);
  pa.mapped.assign(nranks, 0);
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = nranks > 1 ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : CU_MEM_HANDLE_TYPE_NONE;
  size_t gran = 0;
  if (api.GetGran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return "cuMemGetAllocationGranularity failed";
  pa.stride = ((bytes + gran - 1) / gran) * gran;
  pa.active = true;
  std::string err;
  int sock = -1, myfd = -1;
  do {
    if (api.Create(&pa.handles[rank], pa.stride, &prop, 0) != CUDA_SUCCESS) { err = "cuMemCreate failed"; break; }
    if (nranks > 1) {
      if (api.Export(&myfd, pa.handles[rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) { err = "cuMemExportToShareableHandle failed"; break; }
      sock = fd_socket_open(token, rank);
      if (sock < 0) { err = "cannot bind the descriptor-passing socket"; }
    }
  } while (false);
  if (nranks > 1) {
    barrier();                      // every rank has bound its socket (or failed: the exchange below then fails everywhere it matters)
    if (err.empty()) {
      for (int p = 0; p < nranks && err.empty(); ++p)
        if (p != rank && !fd_send(sock, token, p, rank, myfd)) err = "sending the allocation descriptor failed";
      for (int q = 0; q < nranks - 1 && err.empty(); ++q) {
        int from = -1, fd = -1;
        if (!fd_recv(sock, &from, &fd) || from < 0 || from >= nranks || from == rank) { err = "receiving an allocation descriptor failed"; break; }
        if (api.Import(&pa.handles[from], (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) err = "cuMemImportFromShareableHandle failed";
        close(fd);
      }
    }
    barrier();                      // every descriptor has been received: the sockets and the exported descriptor can go
    if (sock >= 0) close(sock);
    if (myfd >= 0) close(myfd);
  }
  if (err.empty()) {
    if (api.Reserve(&pa.base, pa.stride * (size_t)nranks, gran, 0, 0) != CUDA_SUCCESS) err = "cuMemAddressReserve failed";
  }
  for (int p = 0; p < nranks && err.empty(); ++p) {
    if (api.Map(pa.base + (size_t)p * pa.stride, pa.stride, 0, pa.handles[p], 0) != CUDA_SUCCESS) err = "cuMemMap failed";
    else pa.mapped[p] = 1;
  }
  if (err.empty()) {
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (api.SetAccess(pa.base, pa.stride * (size_t)nranks, &acc, 1) != CUDA_SUCCESS) err = "cuMemSetAccess failed (no peer access between the devices?)";
  }
  if (!err.empty()) peer_arena_destroy(pa);
  return err;
}

}  // namespace feastcuda
