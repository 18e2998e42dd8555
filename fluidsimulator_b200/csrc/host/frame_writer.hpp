// frame_writer.hpp — ASCII VTK PolyData frames + PVD series index, byte-identical to the
// reference io/ module (reference io/include/fluid/vtk_writer.h:9-52, io/src/vtk_writer.cpp).
// "io/ output format stays unchanged" (BASELINE.json north_star): ParaView reads both the same.
#pragma once

#include <cstddef>
#include <string>
#include <vector>

namespace fluid {
namespace b200 {

struct FrameView {  // vtk_writer.h:9-15
  const float* pos_x = nullptr;
  const float* pos_y = nullptr;
  const float* pos_z = nullptr;
  std::size_t count = 0;
  double time = 0.0;
};

// <basename>_%06zu.vtp (vtk_writer.cpp:81-87)
std::string frame_filename(const std::string& basename, std::size_t frame_index);
std::string path_join(const std::string& dir, const std::string& file);

// Renders one frame into `out` (no file IO): the exact bytes the reference writes.
void render_frame(const FrameView& frame, std::string& out);

class FrameWriter {
 public:
  explicit FrameWriter(std::string output_dir, std::string basename = "frame");
  bool write(const FrameView& frame, std::size_t frame_index, std::string* out_path = nullptr) const;

 private:
  std::string dir_, base_;
};

class SeriesWriter {  // the .pvd collection (vtk_writer.cpp:100-134)
 public:
  explicit SeriesWriter(std::string output_dir, std::string basename = "series");
  void add(double time, const std::string& relative_path);
  bool write() const;
  std::string path() const;

 private:
  std::string dir_, base_;
  std::vector<std::pair<double, std::string>> entries_;
};

}  // namespace b200
}  // namespace fluid
