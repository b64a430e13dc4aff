// internal: the host fill of the rotated images as a pipeline (host_expand.cpp)
#pragma once
#include <cstdint>

namespace cmg
{
struct ExpandPipeline;
// starts `threads` workers over the packed HOST matrix (dimension 12 nside^2 or 3 x that); directMask bit 3 strip + (k - 1):
// the image in the k-th face below the last one of every ring is not written for that strip (it reaches the host by a direct copy)
ExpandPipeline* expandBegin(double* packed, int64_t nside, int threads, int directMask = 0);
// columns [q0, q1) of the last face of ring `ring` (base face 4 ring + 3) in strip `strip` are complete in host memory:
// their images in the three other faces of the ring may be written now
void expandPublish(ExpandPipeline* p, int strip, int ring, int64_t q0, int64_t q1);
// waits until everything published has been filled in; releases the pipeline
void expandFinish(ExpandPipeline* p);
}
