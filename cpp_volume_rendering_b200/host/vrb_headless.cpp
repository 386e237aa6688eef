// vrb_headless.cpp -- headless driver: the analogue of cppvolrend/main.cpp:49-80 without a window.
//   vrb_headless <data-dir> [--renderer s_1rc] [--size W H] [--camera-state N] [--set Name=value]... [--out frame.f32]
//                [--gradient 0..3] [--screenshot file.png] [--eval DIR [--eval-frames N]]
// --eval runs the reference's evaluation harness (Evaluation.md): DIR/eval_DD-MM-YYYY_HH-MM-SS/{eval.csv, img/NNNN.png}.
// Reads #list_structured_datasets / #list_transfer_functions / #list_camera_states / #list_light_sources from
// <data-dir>, renders one frame with the chosen renderer and writes W*H*4 float32 (row 0 = bottom) to --out.
#include "vrbhost.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

void vrbh_register_renderers(RenderingManager* m);

int main(int argc, char** argv) {
  const char* usage = "usage: %s <data-dir> [--renderer s_1rc|s_1rc_eb|s_1rc_dos|s_1rc_gt_c|s_1rc_vct|iso] [--size W H] [--camera-state N]\n"
                      "          [--device N] [--set Name=value]... [--gradient 0..3] [--out frame.f32] [--screenshot file.png]\n"
                      "          [--eval DIR [--eval-frames N]]\n"
                      "<data-dir> holds #list_structured_datasets, #list_transfer_functions, #list_camera_states, #list_light_sources\n"
                      "(the reference's data/ layout); volumes: .raw, .syn, .pvm (plain or DDS v3d/v3e).  Needs a CUDA device.\n";
  if (argc >= 2 && (std::string(argv[1]) == "--help" || std::string(argv[1]) == "-h")) { printf(usage, argv[0]); return 0; }
  if (argc < 2) { fprintf(stderr, usage, argv[0]); return 2; }
  std::string data = argv[1], renderer = "s_1rc", out, shot, eval_dir;
  int W = 768, H = 768, cam = 0, device = 0, gradient = 3, eval_frames = 100;
  std::vector<std::pair<std::string, double>> sets;
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--renderer" && i + 1 < argc) renderer = argv[++i];
    else if (a == "--size" && i + 2 < argc) { W = atoi(argv[++i]); H = atoi(argv[++i]); }
    else if (a == "--camera-state" && i + 1 < argc) cam = atoi(argv[++i]);
    else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
    else if (a == "--out" && i + 1 < argc) out = argv[++i];
    else if (a == "--screenshot" && i + 1 < argc) shot = argv[++i];
    else if (a == "--gradient" && i + 1 < argc) gradient = atoi(argv[++i]);
    else if (a == "--eval" && i + 1 < argc) eval_dir = argv[++i];
    else if (a == "--eval-frames" && i + 1 < argc) eval_frames = atoi(argv[++i]);
    else if (a == "--set" && i + 1 < argc) {
      std::string kv = argv[++i]; size_t eq = kv.find('=');
      if (eq == std::string::npos) { fprintf(stderr, "--set wants Name=value\n"); return 2; }
      sets.push_back({kv.substr(0, eq), atof(kv.substr(eq + 1).c_str())});
    } else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  RenderingManager* m = RenderingManager::Instance();
  if (!m->InitGL(device)) { fprintf(stderr, "InitGL: %s\n", vrb::LastError().c_str()); return 1; }
  vrbh_register_renderers(m);
  m->GetRenderingParameters()->SetScreenSize(W, H);
  m->GetDataManager()->SetCurrentGradient(gradient);       // generated with the volume (datamanager.cpp:326-327)
  if (!m->InitData(data)) { fprintf(stderr, "InitData: %s\n", vrb::LastError().c_str()); return 1; }
  if (!m->SetCurrentVolumeRendererByAbbreviation(renderer)) { fprintf(stderr, "renderer: %s\n", vrb::LastError().c_str()); return 1; }
  m->Reshape(W, H);
  if (cam > 0 && !m->SetCameraState(cam)) { fprintf(stderr, "camera: %s\n", vrb::LastError().c_str()); return 1; }
  for (auto& kv : sets)
    if (!m->GetCurrentVolumeRenderer()->SetMultiScalingOption(kv.first, kv.second) && !m->GetCurrentVolumeRenderer()->SetParameter(kv.first, kv.second)) {
      fprintf(stderr, "unknown parameter %s\n", kv.first.c_str()); return 1;
    }
  if (!eval_dir.empty()) {
    if (!m->RunEvaluation(eval_dir, eval_frames)) { fprintf(stderr, "evaluation: %s\n", vrb::LastError().c_str()); return 1; }
    printf("evaluation written to %s\n", m->GetEvaluationDirectory().c_str());
  }
  if (!m->Display()) { fprintf(stderr, "Display: %s\n", vrb::LastError().c_str()); return 1; }
  std::vector<float> px;
  if (!m->GetCurrentVolumeRenderer()->ReadOutputRGBA32F(px)) { fprintf(stderr, "read-back: %s\n", vrb::LastError().c_str()); return 1; }
  double sum = 0; size_t hit = 0;
  for (size_t i = 0; i < px.size(); i += 4) { sum += px[i + 3]; hit += px[i + 3] > 0.0f; }
  printf("%s: %dx%d, %zu pixels hit, mean alpha %.6f\n", m->GetCurrentVolumeRenderer()->GetAbbreviationName(), W, H, hit, sum / (px.size() / 4));
  if (!shot.empty() && !m->SaveScreenshot(shot)) { fprintf(stderr, "screenshot: %s\n", vrb::LastError().c_str()); return 1; }
  if (!out.empty()) {
    FILE* f = fopen(out.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
    fwrite(px.data(), sizeof(float), px.size(), f);
    fclose(f);
  }
  RenderingManager::DestroyInstance();
  vrb::Device::Instance()->Shutdown();
  return 0;
}
