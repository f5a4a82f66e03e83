/* placeholders until the dedicated kernels land */
#include "amh_host.h"
namespace amhh {
int launch_mala(amh_run&, int, const amhd::SaveArgs&) { return fail(AMH_ERR_UNSUPPORTED, "MALA kernel not built"); }
int launch_ram(amh_run&, int, bool, const amhd::SaveArgs&) { return fail(AMH_ERR_UNSUPPORTED, "RAM kernel not built"); }
int launch_stretch(amh_run&, int, const amhd::SaveArgs&) { return fail(AMH_ERR_UNSUPPORTED, "stretch kernel not built"); }
}
