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
}  // namespace rlpt
