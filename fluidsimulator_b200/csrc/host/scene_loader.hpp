// scene_loader.hpp — scene JSON -> Params + initial State.
// Same contract as the reference's fluid/init.h:10-27.
#pragma once

#include <cstddef>
#include <string>

#ifdef PBF_USE_REFERENCE_HEADERS
#include "fluid/core.h"
#else
#include "fluid_types.hpp"
#endif

namespace fluid {
namespace b200 {

// Six planes of an axis-aligned box with its near corner at the origin (init.cpp:109-117).
void box_planes(Params& params, float box_x, float box_y, float box_z);
// nx*ny*nz lattice, x fastest (init.cpp:72-107).
void lattice_block(const Params& params, State& state, std::size_t nx, std::size_t ny, std::size_t nz,
                   float origin_x, float origin_y, float origin_z, float spacing = 0.0f);
// The built-in test scene used when no --scene is given (init.cpp:119-156).
void default_test_scene(Params& params, State& state);
// Scene file loader (init.cpp:158-418).  Returns false and fills *error on failure.
bool load_scene_json(const std::string& path, Params& params, State& state, std::string* error = nullptr);

}  // namespace b200
}  // namespace fluid
