/* rpst_debug.h — white-box test hooks of librpst.  NOT part of the product ABI: these entry points exist only in
 * librpst_debug.so (the same sources compiled with -DRPST_DEBUG_EXPORTS; built by rp-style-transfer_b200/build.py
 * next to librpst.so) and are used by tests/test_schedule_gpu.py alone. */
#ifndef RPST_DEBUG_H
#define RPST_DEBUG_H
#include "rpst.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Test hook (no data touched): the ticket schedule the TMA-staged AdaIN kernel walks for a call shape.
 * info[5] (host) = {tickets, statistics items per plane, apply items per plane, lag, merge lead};
 * tickets (device, [max_tickets,3] int32, may be NULL) = (kind, plane, chunk), kind 0 statistics / 1 apply / 2 merge. */
RPST_API int rpst_debug_adain_schedule(int64_t planes, int64_t hw, int has_style, int has_prev, int stats_only,
                              int32_t* tickets, int64_t max_tickets, int64_t* info, void* stream);


/* Test hook: ticket schedule of the segment kernel (see rpst_debug_adain_schedule).  info[5] = {tickets, content
 * statistics items, style statistics items, apply items per plane, lag}; kind 0 content statistics, 1 style
 * statistics, 2 apply, 3 merge. */
RPST_API int rpst_debug_seg_schedule(int64_t n, int64_t c, int64_t hw_c, int64_t hw_s, int has_prev, int32_t* tickets,
                            int64_t max_tickets, int64_t* info, void* stream);


#ifdef __cplusplus
}
#endif
#endif
