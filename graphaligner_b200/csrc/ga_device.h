// Device context: one per GPU.  Owns the replicated graph, constant tables, a stream and grow-only scratch.
#ifndef GA_DEVICE_H
#define GA_DEVICE_H
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include "alignment_graph.h"
#include "ga_host.h"
#include "ga_types.h"

namespace ga
{

// A batch whose inputs already live in HBM (parts + stream descriptors + per-warp layout): the timed region of
// the "inputs resident" benchmark is RunStaged alone.
struct StagedBatch;

DeviceCtx* CreateDevice(int device);
void DestroyDevice(DeviceCtx* ctx);
const std::string& LastError(DeviceCtx* ctx);
void SetError(DeviceCtx* ctx, const std::string& msg);
// copies the finalized graph's flat arrays to the device (replica per GPU)
void UploadGraph(DeviceCtx* ctx, const AlignmentGraph& graph);
size_t GraphBytesOnDevice(DeviceCtx* ctx);
// INT32 lane-ops per second of a dependency-free LOP3/IADD3 kernel on this device (roofline denominator)
double MeasureInt32Peak(DeviceCtx* ctx);

// plans the per-warp memory layout and copies the reads' bytes (`parts`) + stream descriptors to the device.  `parts` must
// outlive the batch.  partsOnDevice: they were already uploaded range by range (UploadPartsRange) while they were staged
StagedBatch* StageAndUpload(DeviceCtx* ctx, const std::vector<ga_stream_in>& streams, const uint8_t* parts, size_t partsBytes, int initialBandwidth, int rampBandwidth, BatchStats* stats,
	bool partsOnDevice = false);
// where each read lies in `parts` (n + 1 offsets): the batch's run then also checks every read for characters the reference
// aborts on; FinishStaged hands the flags back through `badChar` (one byte per read)
void SetReadRanges(DeviceCtx* ctx, StagedBatch* batch, const std::vector<uint64_t>& readOff);
// asynchronous upload of parts[offset, offset + bytes) (parts = the buffer AllocPinnedParts returned) to the same offset on the device
void UploadPartsRange(DeviceCtx* ctx, const uint8_t* parts, size_t offset, size_t bytes);
// pinned host memory for a batch's padded parts (grow-only, owned by the context, reused by the next batch)
uint8_t* AllocPinnedParts(DeviceCtx* ctx, size_t bytes);
// the device twin alone (the reads are uploaded from the caller's own page-locked buffer)
void EnsureDeviceParts(DeviceCtx* ctx, size_t bytes);
// true when `p` lies in page-locked host memory the CUDA runtime knows about (an asynchronous copy from it needs no staging)
bool IsPinnedHost(const void* p);
// sizing helpers for splitting a batch that would not fit the device in one launch
size_t EstimateStreamBytes(DeviceCtx* ctx, size_t partLen, int bandwidth);
size_t FreeDeviceBytes(DeviceCtx* ctx);
// launches the alignment kernel(s) on the context's stream (asynchronous); returns number of launches
int RunStaged(DeviceCtx* ctx, StagedBatch* batch);
// waits, copies results back, re-runs streams that overflowed their scratch with larger capacities
// mapTail (optional): the arena is allocated with room behind the device's records for one 32-byte mapping record per run of
// every stream whose mappings the device did not write (ga_stream_out::nMapped == 0); *mapTail = first word of that room
// (a multiple of 8), arena.size() covers it
void FinishStaged(DeviceCtx* ctx, StagedBatch* batch, RawBuffer<ga_stream_out>& outs, RawBuffer<uint32_t>& arena, BatchStats* stats, std::vector<uint8_t>* badChar = nullptr,
	size_t* mapTail = nullptr);
void FreeStaged(DeviceCtx* ctx, StagedBatch* batch);
void* DeviceStream(DeviceCtx* ctx);   // cudaStream_t
void SyncDevice(DeviceCtx* ctx);

}

#endif
