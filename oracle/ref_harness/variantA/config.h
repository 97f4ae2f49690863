/* TEST INFRASTRUCTURE: strategy switches for the split-enabled oracle ("variant A",
 * SURVEY.md section 8c): the reference's include/config.h with DONT_PARTITION and
 * PRE_PARTITION left undefined, so Shade()'s centre-split path (shade.cpp:463-525)
 * is compiled in.  PARALLEL is NOT defined here because include/config.h:11-13
 * would force DONT_PARTITION; the harness drives parallelRadiosity only when
 * PARALLEL is defined, so variant A defines HARNESS_VARIANT_A and gets PARALLEL
 * through the two lines below without re-enabling DONT_PARTITION. */
#ifndef _CONFIG_H
#define _CONFIG_H
#define ADAPT
#define ROLLBACK
#define PARALLEL
#endif
