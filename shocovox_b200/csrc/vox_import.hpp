// MagicaVoxel `.vox` import (vox_import.cpp): Octree::load_vox_file, reference src/convert/magicavoxel.rs:207-385
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "host_octree.hpp"

namespace svx {

// tree size load_vox_file would pick for this file: next power of two of the largest extent (magicavoxel.rs:266-271)
int32_t vox_required_tree_size(const uint8_t* data, size_t len, uint32_t* tree_size, std::string* why);
// Octree::load_vox_file (:266-289): a new tree of that size holding every voxel of frame 0
int32_t vox_load(const uint8_t* data, size_t len, uint32_t brick_dim, HostOctree** out, std::string* why);
// load_vox_data_internal (:349-385) into an existing tree (e.g. one whose MIP strategy was configured first, which is what
// MIPMapStrategy::load_vox_file does, :207-250)
int32_t vox_insert_into(const uint8_t* data, size_t len, HostOctree* tree, std::string* why);
int32_t vox_read_file(const char* path, std::vector<uint8_t>* bytes);

}  // namespace svx
