/* tests/host_core/context_emu_absent.cpp -- companion of context_emu.cpp (kept apart because these stand-ins deliberately do not match
 * the prototypes of include/cube_slam_b200.h). */
#include <cmath>
#define CS_ERR_UNSUPPORTED (-6)
/* the exports of the translation units that are not part of this build, so that cube_slam_b200/_lib.py can bind the file like the real
 * library; none of them is reached by the tests that use it */
#define CS_EMU_ABSENT(name) \
    extern "C" int name() { return CS_ERR_UNSUPPORTED; }
CS_EMU_ABSENT(cs_detect_lines)
CS_EMU_ABSENT(cs_detect_lines_batch)
CS_EMU_ABSENT(cs_debug_lsd)
CS_EMU_ABSENT(cs_debug_lsd_stats)
CS_EMU_ABSENT(cs_debug_lsd_prof)
CS_EMU_ABSENT(cs_debug_atan2)
CS_EMU_ABSENT(cs_debug_edlines)
CS_EMU_ABSENT(cs_keylines_from_lines)
CS_EMU_ABSENT(cs_lbd_compute)
CS_EMU_ABSENT(cs_lbd_compute_batch)
CS_EMU_ABSENT(cs_detect_descrip_lines)
CS_EMU_ABSENT(cs_detect_descrip_lines_batch)
CS_EMU_ABSENT(cs_match_line_descrip)
CS_EMU_ABSENT(cs_match_line_descrip_batch)
CS_EMU_ABSENT(cs_lbd_debug_prepare)
CS_EMU_ABSENT(cs_lbd_debug_keylines_edl)
extern "C" double cs_atan2_host(double y, double x) { return std::atan2(y, x); }
