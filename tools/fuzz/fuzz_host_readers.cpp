#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
extern "C" {
void* vrbh_volume_read(const char* path);
void vrbh_volume_destroy(void* v);
void* vrbh_tf_read(const char* path);
void vrbh_tf_destroy(void* tf);
int vrbh_read_camera_states(const char* path, float* out9, int cap);
int vrbh_read_light_lists(const char* path, float* out13, int cap);
}
static void put(const char* path, const std::string& s) { FILE* f = fopen(path, "wb"); fwrite(s.data(), 1, s.size(), f); fclose(f); }
static std::string mutate(std::string s, std::mt19937& rng) {
  int k = rng() % 4;
  for (int i = 0; i < k && !s.empty(); ++i) {
    size_t p = rng() % s.size();
    switch (rng() % 4) {
      case 0: s[p] = (char)rng(); break;
      case 1: s.erase(p, rng() % 8); break;
      case 2: s.insert(p, std::to_string((int)rng())); break;
      default: s.resize(p); break;
    }
  }
  return s;
}
int main() {
  std::mt19937 rng(7);
  const std::string tf1d = "1\n3\n0.1 0.2 0.3 0\n0.5 0.5 0.5 100\n0.9 0.8 0.7 255\n2\n0.0 0\n0.8 255\n";
  const std::string tf1d_b = "linear\n0\n2\n0.1 0.2 0.3 0\n0.9 0.8 0.7 255\n2\n0.0 0\n0.8 255\n63\n1\n";
  const std::string syn = "16 16 16\n1 2 2 2 9 9 9 200\n0 3 4 5 77\n1 0 0 0 4 4 4 10\n";
  const std::string cams = "A\nARCBALL\n1 2 3\n0 0 0\n0 1 0\nB\nFLIGHT\nC\nARCBALL\n4 5 6\n0 0 0\n0 0 1";
  const std::string lights = "L\n2\n1 2 3\n0 0 1\n0 1 0\n1 0 0\n5\n4 5 6\n0 0 -1\n0 1 0\n-1 0 0\n7\nM\n1\n1 1 1\n0 0 1\n0 1 0\n1 0 0\n4";
  std::vector<float> out(64 * 13);
  int ok = 0;
  for (int it = 0; it < 4000; ++it) {
    put("/tmp/fz.tf1d", mutate(it & 1 ? tf1d : tf1d_b, rng));
    if (void* t = vrbh_tf_read("/tmp/fz.tf1d")) { vrbh_tf_destroy(t); ++ok; }
    put("/tmp/fz.syn", mutate(syn, rng));
    if (void* v = vrbh_volume_read("/tmp/fz.syn")) { vrbh_volume_destroy(v); ++ok; }
    put("/tmp/fz_cams", mutate(cams, rng));
    vrbh_read_camera_states("/tmp/fz_cams", out.data(), 64);
    put("/tmp/fz_lights", mutate(lights, rng));
    vrbh_read_light_lists("/tmp/fz_lights", out.data(), 64);
    std::string raw((size_t)(rng() % 300), 'x');
    char name[128];
    snprintf(name, sizeof name, "/tmp/fz.%d.%dx%dx%d.raw", (int)(rng() % 4), (int)(rng() % 9) - 1, (int)(rng() % 9), (int)(rng() % 9));
    put(name, raw);
    if (void* v = vrbh_volume_read(name)) { vrbh_volume_destroy(v); ++ok; }
    remove(name);
  }
  printf("ok (%d inputs accepted)\n", ok);
  return 0;
}
