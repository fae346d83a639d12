// group_records.cu -- match records in the grouped form that crosses PCIe.
//
// The reference reports one callback per match, report(ai, bi, start, end) (pbwt.h:208,213), and its reporters emit runs
// of callbacks that share ai, start and end: matchMaximalWithin walks the block above position i and then the block below
// it (pbwtMatch.c:133-134), matchSequencesSweep the block of a query (pbwtMatch.c:385,433-435).  Only bi changes inside such
// a run.  The grouped stream stores a run once -- {ai, start, end, n} followed by its n values of bi -- which is about a
// third of the 16 bytes per record on config 2 (18.5 M records, 95 MB instead of 295 MB), and the host shim expands it while
// it calls report(), record for record in the same order.  Grouping is by equality with the previous record, so the
// expansion is exact whatever produced the records (a reporter whose two blocks start at the same site becomes one group).
#include "group_records.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

namespace pbwtgpu {
namespace {

struct GroupState {
  long long count;  // group starts at or before this record
  long long start;  // index of the latest of them
};
struct GroupCombine {
  __host__ __device__ GroupState operator()(const GroupState &a, const GroupState &b) const {  // a earlier, b later
    return GroupState{a.count + b.count, b.count ? b.start : a.start};
  }
};
__device__ __forceinline__ bool starts_group(const int4 *rec, long long r) {
  if (r == 0) return true;
  const int4 a = rec[r - 1], b = rec[r];
  return a.x != b.x || a.z != b.z || a.w != b.w;  // ai, start, end
}
struct GroupFlag {
  const int4 *rec;
  __host__ __device__ GroupState operator()(long long r) const {
#ifdef __CUDA_ARCH__
    const bool f = starts_group(rec, r);
    return GroupState{f ? 1 : 0, f ? r : -1};
#else
    return GroupState{0, -1};
#endif
  }
};
using FlagIt = cub::TransformInputIterator<GroupState, GroupFlag, cub::CountingInputIterator<long long>>;

__global__ void __launch_bounds__(256) write_groups_kernel(const int4 *rec, const GroupState *gs, long long n, int32_t *out) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const GroupState g = gs[r];
    const int4 m = rec[r];
    out[4 * g.count + r] = m.y;  // bi, behind the headers of all groups up to and including its own
    if (g.start == r) {          // the group's first record writes ai, start, end
      int32_t *h = out + 4 * (g.count - 1) + r;
      h[0] = m.x; h[1] = m.z; h[2] = m.w;
    }
    if (r + 1 == n || gs[r + 1].start == r + 1)  // its last record writes n
      out[4 * (g.count - 1) + g.start + 3] = (int32_t)(r - g.start + 1);
  }
}

}  // namespace

size_t group_records_temp_bytes(long long n) {
  size_t bytes = 0;
  FlagIt it(cub::CountingInputIterator<long long>(0), GroupFlag{nullptr});
  cub::DeviceScan::InclusiveScan(nullptr, bytes, it, (GroupState *)nullptr, GroupCombine(), (int)n);
  return bytes;
}
size_t group_records_state_bytes(long long n) { return sizeof(GroupState) * (size_t)(n > 0 ? n : 1); }

int launch_group_scan(const pbwtgpu_match *rec, long long n, void *state, void *temp, size_t tempBytes, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  FlagIt it(cub::CountingInputIterator<long long>(0), GroupFlag{reinterpret_cast<const int4 *>(rec)});
  return (int)cub::DeviceScan::InclusiveScan(temp, tempBytes, it, static_cast<GroupState *>(state), GroupCombine(), (int)n, stream);
}
const long long *group_scan_total(const void *state, long long n) {  // device address of the number of groups
  return &static_cast<const GroupState *>(state)[n - 1].count;
}
int launch_write_groups(const pbwtgpu_match *rec, const void *state, long long n, int32_t *out, cudaStream_t stream) {
  if (n <= 0) return 0;
  const long long blocks = (n + 255) / 256 < 148LL * 32 ? (n + 255) / 256 : 148LL * 32;
  write_groups_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const int4 *>(rec), static_cast<const GroupState *>(state), n, out);
  return (int)cudaGetLastError();
}

}  // namespace pbwtgpu
