__global__ void kdiv(float* o, const float* a, const float* b){ int i = threadIdx.x; o[i] = __fdiv_rn(a[i], b[i]); }
__global__ void krcp(float* o, const float* a){ int i = threadIdx.x; o[i] = __frcp_rn(a[i]); }
__global__ void kdivc(float* o, const float* a){ int i = threadIdx.x; o[i] = a[i] / 639.0f; }
