/* LD_PRELOAD shim: time() returns the seed the author of src/homing_missile.cpp recommends (homing_missile.cpp:586-588,
 * "1658778374 -- a very nice example of EKF vs EMCE"), so that the unmodified example -- which seeds with time(NULL) -- is
 * repeatable and the reference build and the drop-in build see the same noise realisations. */
#include <time.h>
time_t time(time_t* t) { const time_t v = (time_t)1658778374; if (t) *t = v; return v; }
