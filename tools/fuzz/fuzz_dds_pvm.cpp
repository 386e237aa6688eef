#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
extern "C" {
long long vrbh_dds_decode(const unsigned char* file, unsigned long long n, unsigned char* out, unsigned long long cap);
long long vrbh_dds_encode(const unsigned char* data, unsigned long long n, unsigned skip, unsigned strip, int version, unsigned char* out, unsigned long long cap);
void* vrbh_volume_read(const char* path);
void vrbh_volume_destroy(void* v);
}
int main() {
  std::mt19937 rng(123);
  std::vector<unsigned char> out(1 << 22);
  long long total = 0;
  for (int it = 0; it < 3000; ++it) {
    size_t n = rng() % 600;
    std::vector<unsigned char> f(8 + n);
    memcpy(f.data(), (it & 1) ? "DDS v3d\n" : "DDS v3e\n", 8);
    for (size_t i = 0; i < n; ++i) f[8 + i] = (unsigned char)rng();
    long long r = vrbh_dds_decode(f.data(), f.size(), out.data(), out.size());
    if (r > 0) total += r;
    // truncated / tiny inputs
    for (size_t k = 0; k < 12 && k <= f.size(); ++k) vrbh_dds_decode(f.data(), k, out.data(), out.size());
  }
  // round trips with odd parameters
  for (int it = 0; it < 300; ++it) {
    size_t n = rng() % 5000;
    std::vector<unsigned char> d(n);
    for (auto& b : d) b = (unsigned char)((rng() % 7 == 0) ? rng() : 3);
    unsigned skip = rng() % 6, strip = rng() % 70000;
    std::vector<unsigned char> enc(n * 2 + 64);
    long long e = vrbh_dds_encode(d.data(), n, skip, strip, 1 + (it & 1), enc.data(), enc.size());
    if (e > (long long)enc.size()) { printf("enc too big\n"); return 1; }
    long long r = vrbh_dds_decode(enc.data(), e, out.data(), out.size());
    if (r != (long long)n || memcmp(out.data(), d.data(), n)) { printf("round trip failed it=%d n=%zu r=%lld\n", it, n, r); return 1; }
  }
  // malformed PVM files through the reader
  const char* heads[] = {"PVM\n", "PVM2\n", "PVM3\n", "PVM\n#c\n", "DDS v3d\n", "PVX\n", ""};
  for (int it = 0; it < 2000; ++it) {
    std::string s = heads[rng() % 7];
    int w = (int)(rng() % 9) - 1, h = (int)(rng() % 9) - 1, d = (int)(rng() % 9) - 1;
    char buf[128];
    snprintf(buf, sizeof buf, "%d %d %d\n", w, h, d); s += buf;
    if (rng() % 2) { snprintf(buf, sizeof buf, "%g %g %g\n", (rng() % 5) - 1.0, 1.0, 2.5); s += buf; }
    snprintf(buf, sizeof buf, "%d\n", (int)(rng() % 4)); s += buf;
    size_t pay = rng() % 700;
    for (size_t i = 0; i < pay; ++i) s.push_back((char)rng());
    if (rng() % 5 == 0) s.resize(rng() % (s.size() + 1));
    FILE* f = fopen("/tmp/fuzz_case.pvm", "wb"); fwrite(s.data(), 1, s.size(), f); fclose(f);
    void* v = vrbh_volume_read("/tmp/fuzz_case.pvm");
    if (v) vrbh_volume_destroy(v);
  }
  printf("ok, %lld bytes decoded from random streams\n", total);
  return 0;
}
