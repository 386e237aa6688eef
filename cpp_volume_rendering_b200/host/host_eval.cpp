// host_eval.cpp -- the evaluation harness of the reference: ParameterSpace sweeps, eval.csv,
// one PNG per sample point, CIEDE2000 difference images.  Mirrors cppvolrend/utils/parameterspace.cpp,
// renderingmanager.cpp:174-181,261-317,409-419,476-492,600-716,803-860 and libs/vis_utils/colorutils.cpp:148-311.
// The reference writes PNG through the closed IM 3.12 library (lib/im_3_12); here: zlib deflate + a 60-line PNG framer.
#include "vrbhost.h"
#include <zlib.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <iomanip>
#include <sstream>
#include <filesystem>

static vrb_ctx* CTX() { return vrb::Device::Instance()->ctx(); }

// ------------------------------------------------------------------ ParameterSpace
// Same contract as cppvolrend/utils/parameterspace.{h,cpp} (the sweep visits the cartesian product of the dimensions, the
// LAST dimension varying fastest; IncrEvaluation returns false once the whole space has been visited), written from that
// contract: the sweep is an odometer over the dimensions.
int ParameterSpace::ComputeNumSamplePoints() {
  long long product = m_dimensions.empty() ? 0 : 1;
  for (size_t k = 0; k < m_dimensions.size(); ++k) product *= m_dimensions[k]->NumSteps();
  m_numsamples_cached = (int)product;
  return m_numsamples_cached;
}
bool ParameterSpace::IncrEvaluation() {
  // odometer: advance the fastest wheel; every wheel that runs past its end is rewound and carries into the next slower one
  for (size_t wheel = m_dimensions.size(); wheel-- > 0;) {
    ParameterRangeBase* r = m_dimensions[wheel];
    r->Incr();
    if (!r->End()) return true;
    r->Start();
  }
  return false;                                          // every wheel wrapped: the space is exhausted
}
// Self-test of the sweep (the reference ships one under the same name, parameterspace.cpp:103-149): a 3 x 4 x 2 space must be
// visited in 24 steps, last dimension fastest, every combination exactly once, and the bound variables must carry the values.
bool ParameterSpaceTest() {
  double a = -1.0; int b = -1; double c = -1.0;
  ParameterSpace space;
  space.AddParameterDimension(new ParameterRangeDouble("a", &a, 0.0, 1.0, 0.5));      // 0, 0.5, 1
  space.AddParameterDimension(new ParameterRangeInt("b", &b, 2, 8, 2));               // 2, 4, 6, 8
  space.AddParameterDimension(new ParameterRangeDouble("c", &c, 10.0, 11.0, 1.0));    // 10, 11
  if (space.GetNumSamplePoints() != 3 * 4 * 2) return false;
  bool seen[3][4][2] = {};
  int visited = 0, last_key = -1;
  space.StartEvaluation();
  do {
    const int ia = (int)(a * 2.0 + 0.5), ib = b / 2 - 1, ic = (int)(c - 10.0 + 0.5);
    if (ia < 0 || ia > 2 || ib < 0 || ib > 3 || ic < 0 || ic > 1 || seen[ia][ib][ic]) return false;
    seen[ia][ib][ic] = true;
    const int key = (ia * 4 + ib) * 2 + ic;              // row-major rank: must grow by one per step
    if (key != last_key + 1) return false;
    last_key = key;
    ++visited;
  } while (space.IncrEvaluation());
  return visited == 24;
}

// ------------------------------------------------------------------ PNG (8-bit RGB, no interlace, filter 0)
static void be32(unsigned char* p, uint32_t v) { p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v; }
static bool png_chunk(FILE* f, const char* type, const unsigned char* data, size_t n) {
  unsigned char hdr[8];
  be32(hdr, (uint32_t)n); std::memcpy(hdr + 4, type, 4);
  uint32_t crc = crc32(0L, hdr + 4, 4);
  if (n) crc = crc32(crc, data, (uInt)n);
  unsigned char tail[4]; be32(tail, crc);
  return fwrite(hdr, 1, 8, f) == 8 && (n == 0 || fwrite(data, 1, n, f) == n) && fwrite(tail, 1, 4, f) == 4;
}
bool WritePNG(const std::string& path, int w, int h, const unsigned char* rgb) {
  if (w <= 0 || h <= 0 || !rgb) { vrb::SetError("WritePNG: bad image"); return false; }
  std::vector<unsigned char> raw((size_t)h * (1 + (size_t)w * 3));
  for (int y = 0; y < h; ++y) {
    raw[(size_t)y * (1 + (size_t)w * 3)] = 0;            // filter type None
    std::memcpy(&raw[(size_t)y * (1 + (size_t)w * 3) + 1], rgb + (size_t)y * w * 3, (size_t)w * 3);
  }
  uLongf zn = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(zn);
  if (compress2(z.data(), &zn, raw.data(), (uLong)raw.size(), 6) != Z_OK) { vrb::SetError("WritePNG: deflate failed"); return false; }
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { vrb::SetError("WritePNG: cannot write " + path); return false; }
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  unsigned char ihdr[13];
  be32(ihdr, (uint32_t)w); be32(ihdr + 4, (uint32_t)h);
  ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;      // 8 bit, colour type 2 (RGB)
  bool ok = fwrite(sig, 1, 8, f) == 8 && png_chunk(f, "IHDR", ihdr, 13) && png_chunk(f, "IDAT", z.data(), zn) && png_chunk(f, "IEND", nullptr, 0);
  ok = (fclose(f) == 0) && ok;
  if (!ok) vrb::SetError("WritePNG: write error on " + path);
  return ok;
}

// ------------------------------------------------------------------ screenshots
// The frame is drawn as a textured quad with glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA) (renderingmanager.cpp:103-104)
// over glClearColor(1, 1, 1, 0) (WHITE_BACKGROUND, :40,108-109) into an 8-bit back buffer: the (premultiplied) colour is
// multiplied by alpha once more -- the reference's own look -- and glReadPixels(GL_RGB, GL_UNSIGNED_BYTE) returns
// round(255 * clamp(c)).  glReadPixels rows run bottom-up and IM writes them as such; a PNG stores the top row first.
bool RenderingManager::GetFrontBufferPixelData(std::vector<unsigned char>& rgb, int* w, int* h) {
  if (!curr_vol_renderer) { vrb::SetError("GetFrontBufferPixelData: no renderer"); return false; }
  std::vector<float> px;
  if (!curr_vol_renderer->ReadOutputRGBA32F(px)) return false;
  const int W = curr_rdr_parameters.GetScreenWidth(), H = curr_rdr_parameters.GetScreenHeight();
  if (px.size() != (size_t)W * H * 4) { vrb::SetError("GetFrontBufferPixelData: output is not screen-sized"); return false; }
  rgb.resize((size_t)W * H * 3);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const float* s = &px[((size_t)y * W + x) * 4];
      unsigned char* d = &rgb[((size_t)(H - 1 - y) * W + x) * 3];
      const float a = std::fmin(std::fmax(s[3], 0.0f), 1.0f);                 // fragment outputs are clamped to [0, 1] before blending
      for (int c = 0; c < 3; ++c) {
        const float src = std::fmin(std::fmax(s[c], 0.0f), 1.0f);
        float v = src * a + 1.0f * (1.0f - a);
        v = std::fmin(std::fmax(v, 0.0f), 1.0f);
        d[c] = (unsigned char)std::floor(v * 255.0f + 0.5f);
      }
    }
  if (w) *w = W;
  if (h) *h = H;
  return true;
}
bool RenderingManager::SaveScreenshot(std::string filename) {
  std::vector<unsigned char> rgb; int w = 0, h = 0;
  if (!GetFrontBufferPixelData(rgb, &w, &h)) return false;
  if (filename.empty()) filename = std::string("screenshot_") + (curr_vol_renderer ? curr_vol_renderer->GetAbbreviationName() : "none") + ".png";
  return WritePNG(filename, w, h, rgb.data());
}

// ------------------------------------------------------------------ evaluation
static double NowMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

bool RenderingManager::StartEvaluation(const std::string& base_directory, int frames_per_sample) {
  if (!curr_vol_renderer || !curr_vol_renderer->IsBuilt()) { vrb::SetError("StartEvaluation: no built renderer"); return false; }
  if (m_eval_running) { vrb::SetError("StartEvaluation: an evaluation is already running"); return false; }
  m_eval_numframes = std::max(std::min(frames_per_sample, 500), 1);          // renderingmanager.cpp:815
  curr_vol_renderer->FillParameterSpace(m_eval_paramspace);                   // :539
  auto t = std::time(nullptr);
  auto tm = *std::localtime(&t);
  std::ostringstream oss;
  oss << std::put_time(&tm, "eval_%d-%m-%Y_%H-%M-%S");
  std::string base = base_directory;
  if (!base.empty() && base.back() != '/') base += '/';
  m_eval_basedirectory = base + oss.str();
  m_eval_imgdirectory = m_eval_basedirectory + "/img";
  std::error_code ec;
  std::filesystem::create_directories(m_eval_basedirectory, ec);
  std::filesystem::create_directories(m_eval_imgdirectory, ec);
  if (ec) { vrb::SetError("StartEvaluation: cannot create " + m_eval_imgdirectory); return false; }
  if (m_eval_csvfile.is_open()) m_eval_csvfile.close();
  m_eval_csvfile.open(m_eval_basedirectory + "/eval.csv", std::ios_base::out);
  if (!m_eval_csvfile.is_open()) { vrb::SetError("StartEvaluation: cannot write eval.csv"); return false; }
  m_eval_paramspace.StartEvaluation();
  for (int i = 0; i < m_eval_paramspace.GetNumDimensions(); i++) m_eval_csvfile << m_eval_paramspace.GetDimensionName(i) << ",";
  m_eval_csvfile << "TimePerFrame (ms),FramesPerSecond,ImageFile\n";
  m_eval_running = true;
  m_eval_currframe = 0;
  m_eval_currsample = 0;
  vrb_ctx_synchronize(CTX());
  m_eval_lasttime = NowMs();
  return true;
}

// the block at the end of RenderingManager::Display (:261-317)
void RenderingManager::EvaluationAfterFrame() {
  m_eval_currframe++;
  if (m_eval_currframe < m_eval_numframes) return;
  vrb_ctx_synchronize(CTX());                             // kernels are asynchronous: the frames of this sample have to be finished
  const double currenttime = NowMs();
  const double time_per_frame = (currenttime - m_eval_lasttime) / m_eval_numframes;
  const double frames_per_second = 1000.0 / time_per_frame;
  std::string imagefilename = std::to_string(m_eval_currsample);
  size_t n_zero = 4;
  imagefilename = std::string(n_zero - std::min(n_zero, imagefilename.length()), '0') + imagefilename + ".png";
  SaveScreenshot(m_eval_imgdirectory + "/" + imagefilename);
  for (int i = 0; i < m_eval_paramspace.GetNumDimensions(); i++) m_eval_csvfile << m_eval_paramspace.GetDimensionValue(i) << ",";
  m_eval_csvfile << std::to_string(time_per_frame) << "," << std::to_string(frames_per_second) << "," << "\"" << imagefilename << "\"\n";
  if (m_eval_paramspace.IncrEvaluation()) {
    m_eval_currsample++;
    m_eval_currframe = 0;
    vrb_ctx_synchronize(CTX());
    m_eval_lasttime = NowMs();
  } else {
    m_eval_paramspace.EndEvaluation();                    // restores the parameters the sweep drove
    m_eval_running = false;
    m_eval_csvfile.close();
    curr_vol_renderer->SetOutdated();
  }
}

bool RenderingManager::RunEvaluation(const std::string& base_directory, int frames_per_sample) {
  if (!StartEvaluation(base_directory, frames_per_sample)) return false;
  while (m_eval_running)
    if (!Display()) { m_eval_running = false; m_eval_paramspace.EndEvaluation(); if (m_eval_csvfile.is_open()) m_eval_csvfile.close(); return false; }
  return true;
}

// ------------------------------------------------------------------ CIEDE2000 (libs/vis_utils/colorutils.cpp)
static const double kPi = 3.14159265358979323846264338327950288;      // glm::pi<double>()
static void RGBtoLAB(const double* I_rgb, double* o_lab) {            // ColorSpaces::RGBtoLAB, USE_OPENCV_RGB_TO_LAB_CONVERSION (:148-178)
  double r = I_rgb[0] / 255.0, g = I_rgb[1] / 255.0, b = I_rgb[2] / 255.0;
  r = ((r > 0.04045) ? pow((r + 0.055) / 1.055, 2.4) : (r / 12.92)) * 100.0;
  g = ((g > 0.04045) ? pow((g + 0.055) / 1.055, 2.4) : (g / 12.92)) * 100.0;
  b = ((b > 0.04045) ? pow((b + 0.055) / 1.055, 2.4) : (b / 12.92)) * 100.0;
  double x = (r * 0.412453 + g * 0.357580 + b * 0.180423) / 95.0456;
  double y = (r * 0.212671 + g * 0.715160 + b * 0.072169) / 100.0000;
  double z = (r * 0.019334 + g * 0.119193 + b * 0.950227) / 108.8754;
  double L = ((y > 0.008856) ? 116.0 * cbrt(y) - 16.0 : 903.3 * y);
  x = ((x > 0.008856) ? cbrt(x) : (7.787 * x) + (16.0 / 116.0));
  y = ((y > 0.008856) ? cbrt(y) : (7.787 * y) + (16.0 / 116.0));
  z = ((z > 0.008856) ? cbrt(z) : (7.787 * z) + (16.0 / 116.0));
  o_lab[0] = L; o_lab[1] = 500.0 * (x - y); o_lab[2] = 200.0 * (y - z);
}
static double DegToRad(double x) { return ((x) * (kPi) / 180); }
static double RadToDeg(double x) { return ((x) / (kPi) * 180); }
double Cie2000Comparison(const double* rgb_a, const double* rgb_b) {   // :221-311
  const double eps = 1e-5;
  double A[3], B[3];
  RGBtoLAB(rgb_a, A); RGBtoLAB(rgb_b, B);
  double c1 = sqrt((A[1] * A[1]) + (A[2] * A[2])), c2 = sqrt((B[1] * B[1]) + (B[2] * B[2]));
  double meanC = (c1 + c2) / 2.0;
  double meanC7 = pow(meanC, 7);
  double g = 0.5 * (1 - sqrt(meanC7 / (meanC7 + 6103515625.)));
  double a1p = A[1] * (1.0 + g), a2p = B[1] * (1.0 + g);
  c1 = sqrt((a1p * a1p) + (A[2] * A[2]));
  c2 = sqrt((a2p * a2p) + (B[2] * B[2]));
  double h1 = fmod(atan2(A[2], a1p) + 2 * kPi, 2 * kPi), h2 = fmod(atan2(B[2], a2p) + 2 * kPi, 2 * kPi);
  double deltaL = B[0] - A[0], deltaC = c2 - c1;
  // the reference tests `c1 * c2 < eps` with a plain `if` that the following if/else chain always overrides: no effect
  double deltah;
  if (fabs(h2 - h1) <= kPi) deltah = h2 - h1;
  else if (h2 > h1) deltah = h2 - h1 - 2 * kPi;
  else deltah = h2 - h1 + 2 * kPi;
  double deltaH = 2 * sqrt(c1 * c2) * sin(deltah / 2);
  double meanL = (A[0] + B[0]) / 2;
  meanC = (c1 + c2) / 2.0;
  meanC7 = pow(meanC, 7);
  double meanH;
  if (fabs(h1 - h2) <= kPi + eps) meanH = (h1 + h2) / 2;
  else if (h1 + h2 < 2 * kPi) meanH = (h1 + h2 + 2 * kPi) / 2;
  else meanH = (h1 + h2 - 2 * kPi) / 2;
  double T = 1 - 0.17 * cos(meanH - DegToRad(30)) + 0.24 * cos(2 * meanH) + 0.32 * cos(3 * meanH + DegToRad(6)) - 0.2 * cos(4 * meanH - DegToRad(63));
  double sl = 1 + (0.015 * pow(meanL - 50, 2)) / sqrt(20 + pow(meanL - 50, 2));
  double sc = 1 + 0.045 * meanC;
  double sh = 1 + 0.015 * meanC * T;
  double rc = 2 * sqrt(meanC7 / (meanC7 + 6103515625.));
  double rt = -sin(DegToRad(60 * exp(-pow((RadToDeg(meanH) - 275) / 25, 2)))) * rc;
  double stsrt = fabs(pow(deltaL / sl, 2) + pow(deltaC / sc, 2) + pow(deltaH / sh, 2) + rt * deltaC / sc * deltaH / sh);
  return sqrt(stsrt);
}

// "Set Reference" keeps the float RGBA of the current output (:600-640); "Generate Diff" (:645-716) maps
// clamp(CIEDE2000 / 100, 0, 1) through white (0) -> red (0.4) -> red (1) and writes Image_diff.png
bool RenderingManager::StoreReferenceImage() {
  if (!curr_vol_renderer) { vrb::SetError("StoreReferenceImage: no renderer"); return false; }
  if (!curr_vol_renderer->ReadOutputRGBA32F(s_ref_image)) return false;
  s_ref_w = curr_rdr_parameters.GetScreenWidth(); s_ref_h = curr_rdr_parameters.GetScreenHeight();
  return true;
}
bool RenderingManager::GenerateDiffImage(const std::string& filename, double* max_delta_e) {
  const int o_w = curr_rdr_parameters.GetScreenWidth(), o_h = curr_rdr_parameters.GetScreenHeight();
  if (s_ref_image.empty() || s_ref_w != o_w || s_ref_h != o_h) { vrb::SetError("GenerateDiffImage: no reference image of this size (StoreReferenceImage)"); return false; }
  std::vector<float> cur;
  if (!curr_vol_renderer || !curr_vol_renderer->ReadOutputRGBA32F(cur) || cur.size() != s_ref_image.size()) { vrb::SetError("GenerateDiffImage: cannot read the current frame"); return false; }
  vis::TransferFunction1D tf1d;
  tf1d.AddAlphaControlPoint(vis::TransferControlPoint(1.0, 0));
  tf1d.AddAlphaControlPoint(vis::TransferControlPoint(1.0, 255));
  tf1d.AddRGBControlPoint(vis::TransferControlPoint(1.0, 1.0, 1.0, 0));
  tf1d.AddRGBControlPoint(vis::TransferControlPoint(1.0, 0.0, 0.0, (int)(0.40 * 255.0)));
  tf1d.AddRGBControlPoint(vis::TransferControlPoint(1.0, 0.0, 0.0, (int)(1.00 * 255.0)));
  tf1d.Build();
  std::vector<unsigned char> out((size_t)o_w * o_h * 3);
  double mx = 0.0;
  for (int y = 0; y < o_h; ++y)
    for (int x = 0; x < o_w; ++x) {
      const size_t i = (size_t)y * o_w + x;
      double rgb_ref[3] = {s_ref_image[i * 4 + 0] * 255.0, s_ref_image[i * 4 + 1] * 255.0, s_ref_image[i * 4 + 2] * 255.0};
      double rgb_cur[3] = {cur[i * 4 + 0] * 255.0, cur[i * 4 + 1] * 255.0, cur[i * 4 + 2] * 255.0};
      double diff_lab = Cie2000Comparison(rgb_ref, rgb_cur);
      mx = std::fmax(mx, diff_lab);
      double c_alpha_r = std::fmin(std::fmax(diff_lab / 100.0, 0.0), 1.0);
      vrb::vec4 clr = tf1d.Get(c_alpha_r * 255.0);
      unsigned char* d = &out[((size_t)(o_h - 1 - y) * o_w + x) * 3];
      d[0] = (unsigned char)(clr.x * 255.0); d[1] = (unsigned char)(clr.y * 255.0); d[2] = (unsigned char)(clr.z * 255.0);
    }
  if (max_delta_e) *max_delta_e = mx;
  return WritePNG(filename, o_w, o_h, out.data());
}

// ------------------------------------------------------------------ extern "C" driver surface
extern "C" {
int vrbh_parameter_space_selftest(void) { return ParameterSpaceTest() ? 0 : 1; }
// A sweep over numeric ranges written as CSV rows (one line per sample point, last dimension fastest); kind: 0 float,
// 1 double, 2 int.  Returns the number of points visited; tests compare it with the reference's own classes.
int vrbh_pspace_enumerate(const double* start_end_incr, const int* kind, int ndims, char* out, int cap, int* num_sample_points) {
  float fv[16]; double dv[16]; int iv[16];
  if (ndims < 1 || ndims > 16) return -1;
  ParameterSpace ps;
  for (int i = 0; i < ndims; ++i) {
    const double a = start_end_incr[3 * i], b = start_end_incr[3 * i + 1], c = start_end_incr[3 * i + 2];
    const std::string name = "p" + std::to_string(i);
    if (kind[i] == 0) ps.AddParameterDimension(new ParameterRangeFloat(name, &fv[i], (float)a, (float)b, (float)c));
    else if (kind[i] == 1) ps.AddParameterDimension(new ParameterRangeDouble(name, &dv[i], a, b, c));
    else ps.AddParameterDimension(new ParameterRangeInt(name, &iv[i], (int)a, (int)b, (int)c));
  }
  *num_sample_points = ps.GetNumSamplePoints();
  std::string s;
  int visited = 0;
  ps.StartEvaluation();
  do {
    for (int i = 0; i < ndims; ++i) { s += ps.GetDimensionValue(i); s += (i + 1 < ndims) ? "," : "\n"; }
    ++visited;
  } while (ps.IncrEvaluation() && visited < 100000);
  ps.EndEvaluation();
  if ((int)s.size() + 1 > cap) return -2;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return visited;
}
int vrbh_save_screenshot(const char* path) { return RenderingManager::Instance()->SaveScreenshot(path ? path : "") ? 0 : 1; }
int vrbh_write_png(const char* path, int w, int h, const unsigned char* rgb_top_first) { return WritePNG(path, w, h, rgb_top_first) ? 0 : 1; }
// runs the whole sweep of the current renderer; out_dir receives the eval_... directory
int vrbh_evaluate(const char* base_directory, int frames_per_sample, char* out_dir, int cap) {
  RenderingManager* m = RenderingManager::Instance();
  if (!m->RunEvaluation(base_directory ? base_directory : ".", frames_per_sample)) return 1;
  if (out_dir && cap > 0) { std::strncpy(out_dir, m->GetEvaluationDirectory().c_str(), (size_t)cap - 1); out_dir[cap - 1] = 0; }
  return 0;
}
int vrbh_eval_num_samples(void) {
  RenderingManager* m = RenderingManager::Instance();
  if (!m->GetCurrentVolumeRenderer()) return -1;
  ParameterSpace ps;
  m->GetCurrentVolumeRenderer()->FillParameterSpace(ps);
  return ps.GetNumSamplePoints();
}
double vrbh_cie2000(const double* rgb_a, const double* rgb_b) { return Cie2000Comparison(rgb_a, rgb_b); }
int vrbh_store_reference_image(void) { return RenderingManager::Instance()->StoreReferenceImage() ? 0 : 1; }
int vrbh_generate_diff_image(const char* path, double* max_delta_e) { return RenderingManager::Instance()->GenerateDiffImage(path, max_delta_e) ? 0 : 1; }
}
