// devbuf.h -- small device/pinned-host buffer helpers and the CUDA error check shared by the translation units of
// libpolb200.so (engine.cu: pair style, multi-GPU layer, Ewald; rigid.cu: rigid-body integrator).
#ifndef POLB200_DEVBUF_H
#define POLB200_DEVBUF_H
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace polb200 {

struct CudaError {
  std::string msg;
};

#define CUDA_CHECK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      throw CudaError{std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                      std::to_string(__LINE__) + ")"};                                           \
  } while (0)

template <class T>
struct DBuf {
  T *p = nullptr;
  size_t cap = 0;
  std::vector<void *> *graveyard = nullptr;  // set: replaced allocations are parked here instead of freed
  void drop()
  {
    if (p && graveyard) graveyard->push_back(p);
    else if (p) cudaFree(p);
    p = nullptr;
  }
  void ensure(size_t n, double slack = 1.1)
  {
    if (n <= cap) return;
    drop();
    cap = (size_t)(n * slack) + 64;
    p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
    if (e != cudaSuccess) {
      cap = 0;
      throw CudaError{std::string("cudaMalloc of ") + std::to_string(n * sizeof(T)) + " bytes: " +
                      cudaGetErrorString(e)};
    }
  }
  void release()
  {
    drop();
    cap = 0;
  }
};

template <class T>
struct HPinned {
  T *p = nullptr;
  size_t cap = 0;
  void ensure(size_t n)
  {
    if (n <= cap) return;
    if (p) cudaFreeHost(p);
    cap = n + n / 8 + 64;
    CUDA_CHECK(cudaMallocHost(&p, cap * sizeof(T)));
  }
  void release()
  {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace polb200

#endif
