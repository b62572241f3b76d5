/* oracle/cpu_engine/cpu_engine_shim.h -- TEST / BASELINE INFRASTRUCTURE ONLY (force-included into every translation unit of the
 * Old_CPU_Rendering_Engine baseline build, oracle/cpu_engine/build.sh).
 *
 * The reference's CPU engine does not compile as committed: Shape::intersects(Ray*, Intersection&, int) is pure
 * virtual (C/objects/shape.h:23) while Triangle declares intersects(Ray*, int) (C/objects/triangle.h:27) and defines neither it nor
 * cramer (C/objects/triangle.h:24; C/objects/triangle.cpp has no such definitions), so Surface is abstract. The committed binary
 * (Old_CPU_Rendering_Engine/Build/Monte_Carlo_Raytracer) exports Triangle::intersects(Ray*, Intersection&, int): the declaration
 * lost its middle parameter after that build. This macro gives the two-argument DECLARATION its parameter back while the reference
 * header is compiled where it lies; three-argument uses (the pure virtual, the call sites in C/rays/ray.cpp:18 and
 * C/lights/area_light_plane.cpp:28) pass through unchanged. The two missing DEFINITIONS are in cpu_engine_harness.cpp. */
#ifndef RLPT_CPU_ENGINE_SHIM_H
#define RLPT_CPU_ENGINE_SHIM_H
#define RLPT_CE_PICK(_1, _2, _3, NAME, ...) NAME
#define RLPT_CE_I2(a, b) intersects(a, Intersection& intersection, b)
#define RLPT_CE_I3(a, b, c) intersects(a, b, c)
#define intersects(...) RLPT_CE_PICK(__VA_ARGS__, RLPT_CE_I3, RLPT_CE_I2, RLPT_CE_I2)(__VA_ARGS__)
#endif
