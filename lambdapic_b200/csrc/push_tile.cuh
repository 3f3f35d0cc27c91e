// Constants shared by the tile kernels (push_tile.cu).
#pragma once

constexpr int LIST_WHOLE_STEP = 1 << 30;            // list entry tag: the particle has not been pushed yet (boundary layer)
constexpr int TILE_MIN_CTAS = 2;                    // resident CTAs per SM the particle kernel is compiled for
constexpr int TILE_PERM_SMEM_LIMIT = 57344 * 4;     // bytes of shared-memory histogram a permutation CTA may use (224 KB)
