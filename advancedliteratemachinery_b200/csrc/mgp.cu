#include "mgp.h"
namespace alm {
struct MgpModel { int dummy; };
void mgp_load(Ctx*, const std::map<std::string, HostTensor>&) { throw AlmError{ALM_ERR_UNSUPPORTED, "MGP-STR path not built yet"}; }
void mgp_forward(Ctx*, const float*, int, float*, float*, float*, float*, int32_t*, float*) { throw AlmError{ALM_ERR_UNSUPPORTED, "MGP-STR path not built yet"}; }
void mgp_free(MgpModel* m) { delete m; }
}
