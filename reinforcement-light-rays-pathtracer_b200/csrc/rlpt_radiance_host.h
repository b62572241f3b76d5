// rlpt_radiance_host.h -- host-side radiance-map construction (see rlpt_radiance_host.cpp)
#pragma once
#include <vector>
namespace rlpt {
struct HostVolume { float pos[3]; float nrm[3]; int surface; };
// one element of the reference's flattened kd-tree (G/radiance_volumes/radiance_tree.cuh:19-27)
struct HostTreeElement { int dim; int leaf; unsigned left, right; float data; float pos[3]; float nrm[3]; };
float host_triangle_area(const float* v9);
void host_triangle_normal(const float* v9, float* n3);
void host_build_radiance_map(const float* surface_v, const float* surface_nrm, int n_surfaces, float area_per_sample,
                             std::vector<HostVolume>& volumes, std::vector<HostTreeElement>& tree);

// Nearest-volume candidate cells (rlpt_device.cuh, vcell_find): a fine uniform grid over the surfaces; for every
// (cell, normal class) pair that a surface of that class passes through, the list of volumes that can be the closest
// same-class volume of some point of the cell. Hash table entries are (cell, class, first candidate, groups of 4).
struct HostVCells {
    float ox, oy, oz, h; int nx, ny, nz;
    std::vector<int> table;          // 4 ints per slot; cell = -1 marks an empty slot; size is a power of two
    std::vector<float> cand;         // 4 floats per candidate: position + volume index (as int bits), padded per list to groups of 4
    size_t keys = 0, listed = 0;
    // second level, only for the pairs where a query can fail the first (no same-class volume within accept_r): every
    // same-class volume the reference's kd search can visit from some point of the cell, with its kd-cell bounds.
    std::vector<int> xtable;         // same slot format; the 4th int is the number of candidates
    std::vector<float> xcand;        // 12 floats per candidate: (position, volume index) (lo.xyz, hi.x) (hi.yz, 0, 0)
    size_t xkeys = 0, xlisted = 0;
};
void host_build_vcells(const float* surface_v, const int* surface_class, int n_surfaces, const std::vector<HostVolume>& volumes,
                       const std::vector<int>& volume_class, const std::vector<HostTreeElement>& tree, float cell_h, float accept_r, float within_abs,
                       HostVCells& out);

// Scan units of the brute-force closest hit's conservative pre-test (rlpt_device.cuh, unit_candidates): pairs of primitives
// that form a parallelogram (a triangle and the triangle that completes it across one of its edges). 16 floats per pair
// (v0, e1, e2, n = e1 x e2, pu, pv, ps, 0), followed by one record per unpaired triangle (n_items records in all);
// slot_gid = primitive id per scan slot: the pairs' two triangles, then the rest.
struct HostScanUnits { std::vector<float> scan; std::vector<int> slot_gid; int n_pairs = 0, n_items = 0; float k1 = 0.f, k2 = 0.f, k3 = 0.f, vmax = 0.f; };
void host_build_scan_units(const float* verts9, int n_primitives, HostScanUnits& out);
}  // namespace rlpt
