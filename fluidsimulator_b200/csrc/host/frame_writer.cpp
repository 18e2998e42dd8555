// frame_writer.cpp — see frame_writer.hpp.  The reference streams every float through
// std::ostream with setprecision(9) << fixed (vtk_writer.cpp:49); "%.9f" produces the same
// digits (both round the exact binary value to 9 decimals), and formatting into one buffer is
// ~20x faster than iostream at 78 MB per million-particle frame.
#include "frame_writer.hpp"

#include <cstdio>
#include <filesystem>
#include <fstream>

namespace fluid {
namespace b200 {
namespace {

bool make_dir(const std::string& path) {
  if (path.empty()) return false;
  std::error_code ec;
  std::filesystem::create_directories(path, ec);
  return !ec;
}

bool spill(const std::string& path, const std::string& bytes) {
  std::ofstream out(path, std::ios::out | std::ios::trunc | std::ios::binary);
  if (!out) return false;
  out.write(bytes.data(), static_cast<std::streamsize>(bytes.size()));
  return static_cast<bool>(out);
}

void append_fixed9(std::string& out, double v) {
  char buf[64];
  const int len = std::snprintf(buf, sizeof(buf), "%.9f", v);
  out.append(buf, static_cast<std::size_t>(len));
}

void append_index(std::string& out, std::size_t v) {
  char buf[32];
  const int len = std::snprintf(buf, sizeof(buf), "%zu", v);
  out.append(buf, static_cast<std::size_t>(len));
}

}  // namespace

std::string frame_filename(const std::string& basename, std::size_t frame_index) {
  char buf[32];
  std::snprintf(buf, sizeof(buf), "_%06zu.vtp", frame_index);
  return basename + buf;
}

std::string path_join(const std::string& dir, const std::string& file) {
  if (dir.empty()) return file;
  const char last = dir.back();
  if (last == '/' || last == '\\') return dir + file;
  return dir + '/' + file;
}

void render_frame(const FrameView& f, std::string& out) {  // vtk_writer.cpp:41-70
  out.clear();
  out.reserve(256 + f.count * 80);
  out += "<?xml version=\"1.0\"?>\n";
  out += "<VTKFile type=\"PolyData\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
  out += "  <PolyData>\n";
  out += "    <Piece NumberOfPoints=\"";
  append_index(out, f.count);
  out += "\" NumberOfVerts=\"";
  append_index(out, f.count);
  out += "\" NumberOfLines=\"0\" NumberOfStrips=\"0\" NumberOfPolys=\"0\">\n";
  out += "      <Points>\n";
  out += "        <DataArray type=\"Float32\" NumberOfComponents=\"3\" format=\"ascii\">\n";
  for (std::size_t i = 0; i < f.count; ++i) {
    out += "          ";
    append_fixed9(out, f.pos_x[i]);
    out += ' ';
    append_fixed9(out, f.pos_y[i]);
    out += ' ';
    append_fixed9(out, f.pos_z[i]);
    out += '\n';
  }
  out += "        </DataArray>\n";
  out += "      </Points>\n";
  out += "      <Verts>\n";
  out += "        <DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n";
  for (std::size_t i = 0; i < f.count; ++i) {
    out += "          ";
    append_index(out, i);
    out += '\n';
  }
  out += "        </DataArray>\n";
  out += "        <DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n";
  for (std::size_t i = 0; i < f.count; ++i) {
    out += "          ";
    append_index(out, i + 1);
    out += '\n';
  }
  out += "        </DataArray>\n";
  out += "      </Verts>\n";
  out += "    </Piece>\n";
  out += "  </PolyData>\n";
  out += "</VTKFile>\n";
}

FrameWriter::FrameWriter(std::string output_dir, std::string basename)
    : dir_(std::move(output_dir)), base_(std::move(basename)) {}

bool FrameWriter::write(const FrameView& frame, std::size_t frame_index, std::string* out_path) const {
  if (!frame.pos_x || !frame.pos_y || !frame.pos_z) return false;
  if (!make_dir(dir_)) return false;
  const std::string path = path_join(dir_, frame_filename(base_, frame_index));
  std::string bytes;
  render_frame(frame, bytes);
  if (!spill(path, bytes)) return false;
  if (out_path) *out_path = path;
  return true;
}

SeriesWriter::SeriesWriter(std::string output_dir, std::string basename)
    : dir_(std::move(output_dir)), base_(std::move(basename)) {}

void SeriesWriter::add(double time, const std::string& relative_path) { entries_.emplace_back(time, relative_path); }

std::string SeriesWriter::path() const { return path_join(dir_, base_ + ".pvd"); }

bool SeriesWriter::write() const {  // vtk_writer.cpp:107-130
  if (!make_dir(dir_)) return false;
  std::string out;
  out += "<?xml version=\"1.0\"?>\n";
  out += "<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
  out += "  <Collection>\n";
  for (const auto& e : entries_) {
    out += "    <DataSet timestep=\"";
    append_fixed9(out, e.first);
    out += "\" group=\"\" part=\"0\" file=\"";
    out += e.second;
    out += "\"/>\n";
  }
  out += "  </Collection>\n";
  out += "</VTKFile>\n";
  return spill(path(), out);
}

}  // namespace b200
}  // namespace fluid
