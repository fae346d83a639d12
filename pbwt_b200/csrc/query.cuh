// query.cuh -- launcher interface of the query-vs-panel sweep (matchSequencesSweep, pbwtMatch.c:363-443).
#pragma once
#include "common.cuh"

namespace pbwtgpu {

struct QueryArgs {
  const uint32_t *pcols;  // panel sorted columns y_k  [N][pwpc]
  const uint32_t *qcols;  // query sorted columns      [N][qwpc]
  int pwpc, qwpc;
  int M, Q, N;
  const int32_t *paStart, *qaStart;  // NULL = identity
  pbwtgpu_match *rec;
  long long recCap;
  pbwtgpu_nomatch *nomatch;
  long long nomatchCap;
  unsigned long long *counts;  // [0] records, [1] "no match" events
};

int launch_query_sweep(const QueryArgs &args, cudaStream_t stream);

}  // namespace pbwtgpu
