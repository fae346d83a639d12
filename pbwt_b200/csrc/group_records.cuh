// group_records.cuh -- match records -> grouped stream ({ai, start, end, n} + n x bi per run of records that share ai,
// start and end), see group_records.cu.
#pragma once
#include "common.cuh"

namespace pbwtgpu {

size_t group_records_temp_bytes(long long n);
size_t group_records_state_bytes(long long n);
// inclusive scan over the n records: per record (groups so far, index of the latest group start)
int launch_group_scan(const pbwtgpu_match *rec, long long n, void *state, void *temp, size_t tempBytes, cudaStream_t stream);
const long long *group_scan_total(const void *state, long long n);
// the stream: 4 * groups + n int32
int launch_write_groups(const pbwtgpu_match *rec, const void *state, long long n, int32_t *out, cudaStream_t stream);

}  // namespace pbwtgpu
