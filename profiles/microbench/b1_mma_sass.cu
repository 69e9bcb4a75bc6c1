#include <cstdint>
__global__ void k(const uint32_t* a, const uint32_t* b, int* c, int iters) {
  uint32_t a0=a[threadIdx.x], a1=a[threadIdx.x+32], a2=a[threadIdx.x+64], a3=a[threadIdx.x+96];
  uint32_t b0=b[threadIdx.x], b1=b[threadIdx.x+32];
  int d0=0,d1=0,d2=0,d3=0;
  for (int i=0;i<iters;++i) {
    asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(d0),"+r"(d1),"+r"(d2),"+r"(d3) : "r"(a0),"r"(a1),"r"(a2),"r"(a3),"r"(b0),"r"(b1));
    a0 ^= d0;
  }
  c[threadIdx.x]=d0+d1+d2+d3;
}
__global__ void k8(const uint32_t* a, const uint32_t* b, int* c, int iters) {
  uint32_t a0=a[threadIdx.x], a1=a[threadIdx.x+32], a2=a[threadIdx.x+64], a3=a[threadIdx.x+96];
  uint32_t b0=b[threadIdx.x], b1=b[threadIdx.x+32];
  int d0=0,d1=0,d2=0,d3=0;
  for (int i=0;i<iters;++i) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(d0),"+r"(d1),"+r"(d2),"+r"(d3) : "r"(a0),"r"(a1),"r"(a2),"r"(a3),"r"(b0),"r"(b1));
    a0 ^= d0;
  }
  c[threadIdx.x]=d0+d1+d2+d3;
}
