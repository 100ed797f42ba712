// oracle/shims/callisto_stubs.cpp -- TEST INFRASTRUCTURE ONLY.
// Headless no-op definitions of the Callisto scene-graph calls the reference simulator makes
// (declarations: /root/reference/include/callisto.h:195-843; call sites MAIN:133-142, 218-316,
// 1350-1351, 1424).  CAL_CreateGroup/CAL_CloneGroup hand out fresh ids through their int* argument.
// CAL_AddGroupKeyState records the key-frame positions so a test can read back the trajectory the
// unmodified reference main() produced.
#include "callisto.h"
#include <vector>

static int g_next_id = 1;
struct KeyFrame { int group; float t; float p[3]; };
static std::vector<KeyFrame> g_keys;

extern "C" int lqr_ref_num_keyframes(void) { return (int)g_keys.size(); }
extern "C" void lqr_ref_get_keyframe(int i, int* group, float* t, float* p) {
  *group = g_keys[i].group; *t = g_keys[i].t;
  p[0] = g_keys[i].p[0]; p[1] = g_keys[i].p[1]; p[2] = g_keys[i].p[2];
}
extern "C" void lqr_ref_clear_keyframes(void) { g_keys.clear(); g_next_id = 1; }

int CAL_Initialisation(bool, bool, bool) { return 0; }
int CAL_End() { return 0; }
int CAL_SetViewParams(int, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar) { return 0; }
int CAL_LoadTexture(int, char*, float) { return 0; }
int CAL_CreateGroup(int* groupID, int, bool, char*, bool, int) { if (groupID) *groupID = g_next_id++; return 0; }
int CAL_DestroyGroup(int) { return 0; }
int CAL_SetGroupPosition(int, CAL_scalar, CAL_scalar, CAL_scalar) { return 0; }
int CAL_SetGroupQuaternion(int, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar) { return 0; }
int CAL_SetGroupColor(int, float, float, float, float, bool, int) { return 0; }
int CAL_AddGroupKeyState(int groupID, float time, CAL_scalar* position, CAL_scalar*, CAL_scalar*, bool, int) {
  KeyFrame k; k.group = groupID; k.t = time;
  k.p[0] = position ? (float)position[0] : 0.f; k.p[1] = position ? (float)position[1] : 0.f; k.p[2] = position ? (float)position[2] : 0.f;
  g_keys.push_back(k);
  return 0;
}
int CAL_CloneGroup(int* groupIDNew, int, int, bool, char*, bool) { if (groupIDNew) *groupIDNew = g_next_id++; return 0; }
int CAL_SetObjectOrientation(int, CAL_scalar, CAL_scalar, CAL_scalar) { return 0; }
int CAL_SetObjectScaling(int, CAL_scalar, CAL_scalar, CAL_scalar) { return 0; }
int CAL_SetObjectColor(int, float, float, float, float, int) { return 0; }
int CAL_SetObjectTexture(int, int, float, float, int) { return 0; }
int CAL_CreateBox(int, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, int* objID, char*) { if (objID) *objID = g_next_id++; return 0; }
int CAL_CreateSphere(int, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, int* objID, char*) { if (objID) *objID = g_next_id++; return 0; }
int CAL_CreateCylinder(int, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, CAL_scalar, int* objID, char*) { if (objID) *objID = g_next_id++; return 0; }
int CAL_CreateTriangles(int, int, CAL_scalar*, int* objID, char*) { if (objID) *objID = g_next_id++; return 0; }
int CAL_CreatePolygon(int, int, CAL_scalar*, int* objID, char*) { if (objID) *objID = g_next_id++; return 0; }
