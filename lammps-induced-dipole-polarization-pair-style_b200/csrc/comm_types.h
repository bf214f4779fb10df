// comm_types.h -- state of the multi-GPU layer of one pair-style instance (one process per GPU).
//
// NCCL is loaded at run time (dlopen) so that libpolb200.so has no link-time dependency on it: a
// single-GPU caller never touches it, and a torch process that already carries its own libnccl.so.2
// shares that copy instead of loading a second one.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <string>

#include "decomp.h"

namespace polb200 {

struct NcclApi {
  void *lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;

  bool load(std::string &err)
  {
    if (lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = std::string("cannot load NCCL (libnccl.so.2): ") + dlerror();
      return false;
    }
#define POLB200_SYM(field, name)                                   \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, name));     \
  if (!field) {                                                    \
    err = std::string("NCCL symbol missing: ") + name;             \
    lib = nullptr;                                                 \
    return false;                                                  \
  }
    POLB200_SYM(GetUniqueId, "ncclGetUniqueId")
    POLB200_SYM(CommInitRank, "ncclCommInitRank")
    POLB200_SYM(CommDestroy, "ncclCommDestroy")
    POLB200_SYM(Send, "ncclSend")
    POLB200_SYM(Recv, "ncclRecv")
    POLB200_SYM(GroupStart, "ncclGroupStart")
    POLB200_SYM(GroupEnd, "ncclGroupEnd")
    POLB200_SYM(AllReduce, "ncclAllReduce")
    POLB200_SYM(AllGather, "ncclAllGather")
    POLB200_SYM(GetErrorString, "ncclGetErrorString")
#undef POLB200_SYM
    return true;
  }
};

// what the pack kernels need to know about the brick (passed by value)
struct SendGeom {
  double lo[3], hi[3];      // sub-domain bounds
  double cut;               // ghost cutoff = cutneighmax
  unsigned valid;           // bit d: direction d has a destination
  double shift[NDIR][3];    // coordinate shift added by the sender (wrap * box length)
  int code[NDIR];           // packed image code (sx+1) | (sy+1)<<2 | (sz+1)<<4 of that shift
};

// peer-memory view for the fused sweep + halo push (filled when every peer buffer could be mapped)
constexpr int MAX_PEERS = 8;
struct PeerPush {
  int enabled;
  double4 *mu[2][MAX_PEERS];          // [buffer parity][rank]: base of that rank's dipole array (ext order)
  double4 *xq[MAX_PEERS];             // [rank]: base of that rank's position+charge array (ext order)
  unsigned long long *flag[MAX_PEERS];  // [rank]: base of that rank's arrival counters (one per source rank)
};
constexpr int NPEERBUF = 4;  // mapped buffers per rank: mua, mub, arrival counters, xq

}  // namespace polb200
