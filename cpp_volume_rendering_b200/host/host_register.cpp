// host_register.cpp -- renderer registration, the analogue of cppvolrend/main.cpp:57-68
// (RenderingManager::Instance()->AddVolumeRenderer(new X())).
#include "vrbhost.h"

void vrbh_register_renderers(RenderingManager* m) {
  m->AddVolumeRenderer(new RayCasting1Pass());
  m->AddVolumeRenderer(new RayCasting1PassIsoAdapt());
  m->AddVolumeRenderer(new RC1PConeLightGroundTruthSteps());
  m->AddVolumeRenderer(new RC1PConeTracingDirOcclusionShading());
  m->AddVolumeRenderer(new RC1PExtinctionBasedShading());
  m->AddVolumeRenderer(new RC1PVoxelConeTracingSGPU());
}
