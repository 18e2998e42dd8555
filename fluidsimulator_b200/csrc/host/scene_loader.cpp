// scene_loader.cpp — host-side scene initialisation, bit-identical to the reference loader
// (reference core/src/init.cpp).  The initial particle positions define every parity
// comparison, so the float arithmetic follows the reference exactly:
//   spacing = cbrtf(mass / density)                      init.cpp:224-226
//   for (float x = o + spacing*0.5f; x < o + size; x += spacing)   float accumulation, :298-300
//   plane: n / |n|, d = n . point                        init.cpp:403-413
// tests/test_host_tools.py checks it against the reference loader and the golden digests.
#include "scene_loader.hpp"

#include <cmath>
#include <fstream>
#include <sstream>

#include "mini_json.hpp"

namespace fluid {
namespace b200 {
namespace {

using minijson::Value;

void report(std::string* error, const std::string& message) {
  if (error) *error = message;
}

struct Vec3f {
  float x = 0.0f, y = 0.0f, z = 0.0f;
};

bool vec3(const Value& v, Vec3f& out, std::string* error, const char* label) {
  if (!v.is_array() || v.array.size() != 3) {
    report(error, std::string("Expected vec3 for ") + label);
    return false;
  }
  try {
    out.x = v.array[0].as_float();
    out.y = v.array[1].as_float();
    out.z = v.array[2].as_float();
  } catch (const std::exception& ex) {
    report(error, std::string("Invalid vec3 for ") + label + ": " + ex.what());
    return false;
  }
  return true;
}

float lattice_spacing(const Params& params) {  // init.cpp:15-21
  if (params.density <= 0.0f || params.particle_mass <= 0.0f) return 0.0f;
  return std::cbrt(params.particle_mass / params.density);
}

std::size_t cells_along(float extent, float spacing) {  // init.cpp:61-68
  if (extent <= 0.0f || spacing <= 0.0f) return 0;
  const std::size_t count = static_cast<std::size_t>(std::floor(extent / spacing));
  return count > 0 ? count : 1;
}

struct Emitter {
  std::vector<float> px, py, pz, vx, vy, vz;
  void push(float x, float y, float z, const Vec3f& v) {
    px.push_back(x);
    py.push_back(y);
    pz.push_back(z);
    vx.push_back(v.x);
    vy.push_back(v.y);
    vz.push_back(v.z);
  }
};

}  // namespace

void box_planes(Params& params, float box_x, float box_y, float box_z) {
  params.planes.clear();
  params.planes.add_normalized(1.0f, 0.0f, 0.0f, 0.0f);
  params.planes.add_normalized(-1.0f, 0.0f, 0.0f, -box_x);
  params.planes.add_normalized(0.0f, 1.0f, 0.0f, 0.0f);
  params.planes.add_normalized(0.0f, -1.0f, 0.0f, -box_y);
  params.planes.add_normalized(0.0f, 0.0f, 1.0f, 0.0f);
  params.planes.add_normalized(0.0f, 0.0f, -1.0f, -box_z);
}

void lattice_block(const Params& params, State& state, std::size_t nx, std::size_t ny, std::size_t nz,
                   float origin_x, float origin_y, float origin_z, float spacing) {
  const std::size_t count = nx * ny * nz;
  for (std::vector<float>* v : {&state.pos_x, &state.pos_y, &state.pos_z, &state.vel_x, &state.vel_y, &state.vel_z})
    v->assign(count, 0.0f);
  state.time = 0.0f;
  const float step = spacing > 0.0f ? spacing : lattice_spacing(params);
  std::size_t at = 0;
  for (std::size_t k = 0; k < nz; ++k)
    for (std::size_t j = 0; j < ny; ++j)
      for (std::size_t i = 0; i < nx; ++i, ++at) {
        state.pos_x[at] = origin_x + static_cast<float>(i) * step;
        state.pos_y[at] = origin_y + static_cast<float>(j) * step;
        state.pos_z[at] = origin_z + static_cast<float>(k) * step;
      }
}

void default_test_scene(Params& params, State& state) {
  const float box_x = 1.0f, box_y = 3.0f, box_z = 1.0f;
  box_planes(params, box_x, box_y, box_z);
  float spacing = params.particle_radius > 0.0f ? params.particle_radius * 2.0f : 0.0f;
  if (spacing <= 0.0f) spacing = lattice_spacing(params);
  if (spacing <= 0.0f) spacing = 0.02f;
  if (params.density > 0.0f) params.particle_mass = params.density * spacing * spacing * spacing;
  params.h = 2.5f * spacing;
  const std::size_t nx = cells_along(box_x * 0.5f, spacing);
  const std::size_t ny = cells_along(box_y * 0.5f, spacing);
  const std::size_t nz = cells_along(box_z * 0.5f, spacing);
  const float span_x = static_cast<float>(nx) * spacing;
  const float span_z = static_cast<float>(nz) * spacing;
  lattice_block(params, state, nx, ny, nz, 0.5f * (box_x - span_x), 0.5f * box_y, 0.5f * (box_z - span_z), spacing);
}

bool load_scene_json(const std::string& path, Params& params, State& state, std::string* error) {
  std::ifstream in(path);
  if (!in) {
    report(error, "Failed to open scene file: " + path);
    return false;
  }
  std::stringstream buffer;
  buffer << in.rdbuf();
  Value root;
  try {
    root = minijson::parse(buffer.str());
  } catch (const std::exception& ex) {
    report(error, std::string("Failed to parse JSON: ") + ex.what());
    return false;
  }
  const Value* fluid = root.find("fluid");
  if (!fluid || !fluid->is_object()) {
    report(error, "Scene JSON missing fluid object.");
    return false;
  }
  try {  // init.cpp:183-203
    if (const Value* v = fluid->find("particle_mass")) params.particle_mass = v->as_float();
    if (const Value* v = fluid->find("density")) params.density = v->as_float();
    if (const Value* v = fluid->find("h")) params.h = v->as_float();
    if (const Value* v = fluid->find("epsilon")) params.epsilon = v->as_float();
    if (const Value* v = fluid->find("n")) params.scorr_n = v->as_int();
    if (const Value* v = fluid->find("k")) params.scorr_k = v->as_float();
    if (const Value* v = fluid->find("c")) params.visc_c = v->as_float();
  } catch (const std::exception& ex) {
    report(error, std::string("Invalid fluid parameters: ") + ex.what());
    return false;
  }
  if (const Value* forces = root.find("external_forces")) {  // init.cpp:209-220
    Vec3f f;
    if (!vec3(*forces, f, error, "external_forces")) return false;
    params.external_forces.x = f.x;
    params.external_forces.y = f.y;
    params.external_forces.z = f.z;
  }

  // init.cpp:222-242
  float spacing = 0.0f;
  bool from_mass = false;
  if (params.density > 0.0f && params.particle_mass > 0.0f) {
    spacing = std::cbrt(params.particle_mass / params.density);
    from_mass = true;
  }
  if (spacing <= 0.0f && params.particle_radius > 0.0f) spacing = params.particle_radius * 2.0f;
  if (spacing <= 0.0f) spacing = 0.02f;
  if (from_mass || params.particle_radius <= 0.0f) params.particle_radius = spacing * 0.5f;
  if (params.h <= 0.0f) params.h = 2.5f * spacing;

  const Value* shapes = fluid->find("shape");
  if (!shapes || !shapes->is_array()) {
    report(error, "Fluid shape must be an array.");
    return false;
  }
  Emitter out;
  const float half = spacing * 0.5f;
  for (const Value& shape : shapes->array) {  // init.cpp:256-353
    if (!shape.is_object()) {
      report(error, "Fluid shape entry must be an object.");
      return false;
    }
    const Value* type_v = shape.find("type");
    if (!type_v) {
      report(error, "Fluid shape entry missing type.");
      return false;
    }
    std::string type;
    try {
      type = type_v->as_string();
    } catch (const std::exception& ex) {
      report(error, std::string("Invalid shape type: ") + ex.what());
      return false;
    }
    Vec3f vel;
    if (const Value* v = shape.find("velocity"))
      if (!vec3(*v, vel, error, "shape.velocity")) return false;
    if (type == "cube") {
      const Value* o_v = shape.find("origin");
      const Value* s_v = shape.find("size");
      if (!o_v || !s_v) {
        report(error, "Cube shape missing origin or size.");
        return false;
      }
      Vec3f o, s;
      if (!vec3(*o_v, o, error, "shape.origin") || !vec3(*s_v, s, error, "shape.size")) return false;
      for (float x = o.x + half; x < o.x + s.x; x += spacing)
        for (float y = o.y + half; y < o.y + s.y; y += spacing)
          for (float z = o.z + half; z < o.z + s.z; z += spacing) out.push(x, y, z, vel);
    } else if (type == "sphere") {
      const Value* o_v = shape.find("origin");
      const Value* r_v = shape.find("radius");
      if (!o_v || !r_v) {
        report(error, "Sphere shape missing origin or radius.");
        return false;
      }
      Vec3f o;
      if (!vec3(*o_v, o, error, "shape.origin")) return false;
      float radius = 0.0f;
      try {
        radius = r_v->as_float();
      } catch (const std::exception& ex) {
        report(error, std::string("Invalid sphere radius: ") + ex.what());
        return false;
      }
      const float r2 = radius * radius;
      for (float x = o.x - radius + half; x < o.x + radius; x += spacing)
        for (float y = o.y - radius + half; y < o.y + radius; y += spacing)
          for (float z = o.z - radius + half; z < o.z + radius; z += spacing) {
            const float dx = x - o.x, dy = y - o.y, dz = z - o.z;
            if (dx * dx + dy * dy + dz * dz < r2) out.push(x, y, z, vel);
          }
    } else {
      report(error, "Unsupported shape type: " + type);
      return false;
    }
  }
  state.pos_x = std::move(out.px);
  state.pos_y = std::move(out.py);
  state.pos_z = std::move(out.pz);
  state.vel_x = std::move(out.vx);
  state.vel_y = std::move(out.vy);
  state.vel_z = std::move(out.vz);
  state.time = 0.0f;

  params.planes.clear();  // init.cpp:364-415 (the per-plane "friction" entry is never read)
  if (const Value* collisions = root.find("collisions")) {
    if (!collisions->is_array()) {
      report(error, "collisions must be an array.");
      return false;
    }
    for (const Value& entry : collisions->array) {
      const Value* type_v = entry.is_object() ? entry.find("type") : nullptr;
      if (!type_v) {
        report(error, "Collision entry missing type.");
        return false;
      }
      std::string type;
      try {
        type = type_v->as_string();
      } catch (const std::exception& ex) {
        report(error, std::string("Invalid collision type: ") + ex.what());
        return false;
      }
      if (type != "plane") {
        report(error, "Unsupported collision type: " + type);
        return false;
      }
      const Value* p_v = entry.find("point");
      const Value* n_v = entry.find("normal");
      if (!p_v || !n_v) {
        report(error, "Plane collision missing point or normal.");
        return false;
      }
      Vec3f p, n;
      if (!vec3(*p_v, p, error, "collision.point") || !vec3(*n_v, n, error, "collision.normal")) return false;
      const float len_sq = n.x * n.x + n.y * n.y + n.z * n.z;
      if (len_sq <= 0.0f) {
        report(error, "Collision normal must be non-zero.");
        return false;
      }
      const float inv_len = 1.0f / std::sqrt(len_sq);
      const float ux = n.x * inv_len, uy = n.y * inv_len, uz = n.z * inv_len;
      params.planes.add(ux, uy, uz, ux * p.x + uy * p.y + uz * p.z);
    }
  }
  return true;
}

}  // namespace b200
}  // namespace fluid
