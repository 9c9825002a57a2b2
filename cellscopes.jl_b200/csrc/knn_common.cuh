// Per-row running top-k shared by the kNN kernels: a sorted (descending score, ascending id) list
// stored with a stride so that one thread owns one list and the warp's lists interleave bank-free.
#pragma once

// Insert (s, id) into the list whose element j lives at ls[j*stride], li[j*stride]; the caller has
// already checked that the candidate beats the current k-th entry.
__device__ __forceinline__ void topk_insert(float* ls, int* li, int k, int stride, float s, int id) {
  int pos = k - 1;
  while (pos > 0) {
    const float ps = ls[(pos - 1) * stride];
    const int pi = li[(pos - 1) * stride];
    if (s > ps || (s == ps && id < pi)) {
      ls[pos * stride] = ps;
      li[pos * stride] = pi;
      --pos;
    } else {
      break;
    }
  }
  ls[pos * stride] = s;
  li[pos * stride] = id;
}
