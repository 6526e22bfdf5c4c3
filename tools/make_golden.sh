#!/bin/sh
# Regenerates the golden dumps under tests/golden/ by running the UNMODIFIED reference (oracle/_ref/ref_run_cpu1,
# built from /root/reference by `make -C oracle ref`) on every committed scenario.  Full term/table dumps are kept
# for the first few steps only (small files); later steps carry per-shape counts, key digests and moments.
set -e
cd "$(dirname "$0")/.."
python tools/gen_scenarios.py tests/golden
[ -f tests/golden/leo7.mces ] || oracle/_ref/ref_gen_leo7 tests/golden/leo7.mces 4
[ -f tests/golden/leo5.mces ] || oracle/_ref/ref_gen_leo5 tests/golden/leo5.mces 7
R=oracle/_ref/ref_run_cpu1
$R tests/golden/lti3.mces         tests/golden/lti3.ref.mced         --full-upto 5
$R tests/golden/lti2.mces         tests/golden/lti2.ref.mced         --full-upto 6
$R tests/golden/lti4.mces         tests/golden/lti4.ref.mced         --full-upto 4
$R tests/golden/lti3_3msmts.mces  tests/golden/lti3_3msmts.ref.mced  --full-upto 7
$R tests/golden/lti4_2pnoise.mces tests/golden/lti4_2pnoise.ref.mced --full-upto 3
$R tests/golden/lti4_2msmts.mces  tests/golden/lti4_2msmts.ref.mced  --full-upto 5
for n in 2 3 4 5 6 7 8; do $R tests/golden/syn$n.mces tests/golden/syn$n.ref.mced --full-upto 3; done
$R tests/golden/leo7.mces         tests/golden/leo7.ref.mced         --full-upto 3
$R tests/golden/homing3.mces      tests/golden/homing3.ref.mced      --full-upto 5    # control input (B, u) + own-mean re-centring
$R tests/golden/leo5.mces         tests/golden/leo5.ref.mced         --full-upto 4 --max-steps 13
# BASELINE.json configs[1]: the reference's homing-missile EMCE (closed loop, author's seed) recorded open-loop by oracle/ref_gen_homing.cpp
[ -f tests/golden/homing_real.mces ] || oracle/_ref/ref_gen_homing tests/golden/homing_real.mces 8
$R tests/golden/homing_real.mces  tests/golden/homing_real.ref.mced  --full-upto 5
python tools/gen_scenarios.py tests/golden      # again: the deep-declared variants of leo5 need leo5.mces
$R tests/golden/lti3_deep.mces         tests/golden/lti3_deep.ref.mced         --full-upto 4    # max_shape 22
$R tests/golden/lti4_2pnoise_deep.mces tests/golden/lti4_2pnoise_deep.ref.mced --full-upto 3    # max_shape 18
$R tests/golden/lti3_3msmts_deep.mces  tests/golden/lti3_3msmts_deep.ref.mced  --full-upto 5    # max_shape 18
$R tests/golden/leo7_deep16.mces       tests/golden/leo7_deep16.ref.mced       --full-upto 3    # d = 7, max_shape 16
# the example's sliding-window depth (5 time steps = 15 MUs): recorded with foo_steps = 5, golden for MUs 1..13 (19 minutes on one thread)
[ -f tests/golden/leo7_w5.mces ] || oracle/_ref/ref_gen_leo7 tests/golden/leo7_w5.mces 5
[ -f tests/golden/leo7_w5.ref.mced ] || $R tests/golden/leo7_w5.mces tests/golden/leo7_w5.ref.mced --full-upto 0 --max-steps 13
# the 8-thread reference (shipping default NUM_CPUS=8), informational: counts and moments only
oracle/_ref/ref_run_cpu8 tests/golden/leo7.mces tests/golden/leo7.ref8.mced --full-upto 0
ls -la tests/golden
# point-wise 1-D marginal cpdf of the reference (cpdf_ndim.hpp:1233-1354, 2055-2139) after selected steps, every state index,
# and its 2-D marginal (cpdf_ndim.hpp:1356-1455, 1774-1919) of the state pairs (0,1), (1,2), ..., (0,d-1) on a 21 x 21 grid
C=oracle/_ref/ref_cpdf_cpu1
$C tests/golden/lti3.mces         tests/golden/lti3.cpdf.mced         -2.0 2.0 0.05  2,5,8,10 --2d -1.0 1.0 0.1 -1.5 1.5 0.15
$C tests/golden/lti4_2pnoise.mces tests/golden/lti4_2pnoise.cpdf.mced -3.0 3.0 0.1   3,6 --2d -2.0 2.0 0.2 -2.0 2.0 0.2
$C tests/golden/syn5.mces         tests/golden/syn5.cpdf.mced         -1.0 1.0 0.04  4,6 --2d -1.0 1.0 0.1 -1.0 1.0 0.1
$C tests/golden/leo5.mces         tests/golden/leo5.cpdf.mced         -0.5 0.5 0.01  5,10 --2d -0.3 0.3 0.03 -0.3 0.3 0.03
$C tests/golden/leo7.mces         tests/golden/leo7.cpdf.mced         -0.2 0.2 0.005 6,11 --2d -0.1 0.1 0.01 -0.1 0.1 0.01
$C tests/golden/lti2.mces         tests/golden/lti2.cpdf.mced         -2.0 2.0 0.05  3,7,9 --2d -2.0 2.0 0.2 -2.0 2.0 0.2
$C tests/golden/lti3_3msmts.mces  tests/golden/lti3_3msmts.cpdf.mced  -2.0 2.0 0.05  4,8,12 --2d -2.0 2.0 0.2 -2.0 2.0 0.2
$C tests/golden/lti4_2msmts.mces  tests/golden/lti4_2msmts.cpdf.mced  -2.0 2.0 0.05  3,7,9 --2d -2.0 2.0 0.2 -2.0 2.0 0.2
$C tests/golden/syn8.mces         tests/golden/syn8.cpdf.mced         -1.0 1.0 0.04  2,4 --2d -1.0 1.0 0.1 -1.0 1.0 0.1
$C tests/golden/homing3.mces      tests/golden/homing3.cpdf.mced      -1.0 1.0 0.02  3,6 --2d -1.0 1.0 0.1 -1.0 1.0 0.1
# deterministic_time_prop / shift_cf_by_bias golden vectors (oracle/ref_transforms.cpp)
oracle/_ref/ref_transforms_cpu1 tests/golden/lti3.mces    tests/golden/lti3.transforms.mced    4
oracle/_ref/ref_transforms_cpu1 tests/golden/homing3.mces tests/golden/homing3.transforms.mced 4
