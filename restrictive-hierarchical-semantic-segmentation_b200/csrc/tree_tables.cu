// Class tree -> per-level index tables (host side) and library-level helpers.
// Replaces the name-keyed structures of Models/models.py:38-54, :82-98, :229-238 and the
// per-forward string lookups (:293, :789; Metrics/losses.py:168).
#include "common.cuh"

extern "C" int rhseg_abi_version(void) { return RHSEG_ABI_VERSION; }

extern "C" const char* rhseg_status_string(int status) {
  switch (status) {
    case RHSEG_OK: return "ok";
    case RHSEG_ERR_ARG: return "rhseg: invalid argument (null pointer, non-positive size or bad stride)";
    case RHSEG_ERR_UNSUPPORTED: return "rhseg: unsupported size (channels per level outside the compiled range)";
    case RHSEG_ERR_TREE: return "rhseg: malformed level description (children of a parent must be contiguous)";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "rhseg: unknown status";
}

extern "C" int rhseg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  RHSEG_CUDA(cudaGetDevice(&dev));
  if (sm_count) RHSEG_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) RHSEG_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) RHSEG_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return RHSEG_OK;
}

extern "C" int rhseg_tree_compile_level(const int32_t* parent_ch, int K, int K_prev, int32_t* table) {
  if (!parent_ch || !table || K < 1) return RHSEG_ERR_ARG;
  if (K > RHSEG_MAX_K || K_prev > RHSEG_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  for (int i = 0; i < RHSEG_TABLE_INTS; ++i) table[i] = -1;
  table[0] = K;
  table[2] = K_prev;
  const bool root = parent_ch[0] < 0;
  int G = 0;
  for (int k = 0; k < K; ++k) {
    const int p = parent_ch[k];
    if (root) {
      if (p >= 0) return RHSEG_ERR_TREE;  // mixed root / non-root channels in one level
      table[RHSEG_TBL_PARENT + k] = -1;
      table[RHSEG_TBL_GROUP_OF + k] = k;  // every root channel is its own (sigmoid) group
      table[RHSEG_TBL_GSTART + k] = k;
      table[RHSEG_TBL_GLEN + k] = 1;
      G = K;
      continue;
    }
    if (p < 0 || p >= K_prev) return RHSEG_ERR_TREE;
    if (k == 0 || p != parent_ch[k - 1]) {
      for (int g = 0; g < G; ++g)
        if (table[RHSEG_TBL_GPARENT + g] == p) return RHSEG_ERR_TREE;  // parent seen before: not contiguous
      table[RHSEG_TBL_GSTART + G] = k;
      table[RHSEG_TBL_GLEN + G] = 0;
      table[RHSEG_TBL_GPARENT + G] = p;
      ++G;
    }
    table[RHSEG_TBL_PARENT + k] = p;
    table[RHSEG_TBL_GROUP_OF + k] = G - 1;
    table[RHSEG_TBL_GLEN + G - 1] += 1;
  }
  table[1] = G;
  table[3] = root ? RHSEG_ACT_SIGMOID : RHSEG_ACT_GROUPED;
  return RHSEG_OK;
}
