#!/bin/bash
# SASS mnemonics that prove which hardware paths the kernels use (run in the build container: cuobjdump on the objects).
L=xtensor_b200/lib
count() { cuobjdump -sass $L/$1 2>/dev/null | grep -c "$2"; }
echo "# sm_100a SASS of libxtb200's objects (cuobjdump -sass), occurrences per object"
echo "xtb_static_f64.o  UTMALDG (TMA tile loads, k_ew_tile_tma)          : $(count xtb_static_f64.o UTMALDG)"
echo "xtb_static_f64.o  UTMASTG (TMA tile stores)                        : $(count xtb_static_f64.o UTMASTG)"
echo "xtb_static_f32.o  UTMALDG                                          : $(count xtb_static_f32.o UTMALDG)"
echo "xtb_scan.o        UTMALDG|UTMASTG (k_scan_colwalk / coltile boxes) : $(count xtb_scan.o 'UTMALDG\|UTMASTG')"
echo "xtb_scan.o        UBLKCP (cp.async.bulk, k_scan_stile / ahead)     : $(count xtb_scan.o UBLKCP)"
echo "xtb_scan.o        LDGSTS (cp.async, k_scan_colwalk_plain)          : $(count xtb_scan.o LDGSTS)"
echo "xtb_scan.o        SYNCS (mbarrier)                                 : $(count xtb_scan.o SYNCS)"
echo "xtb_static_reduce.o ACQBULK (griddepcontrol.wait, PDL)             : $(count xtb_static_reduce.o ACQBULK)"
echo "xtb_static_reduce.o PREEXIT (griddepcontrol.launch_dependents)     : $(count xtb_static_reduce.o PREEXIT)"
echo "xtb_static_reduce.o LDG.E.NA.128.CONSTANT (streaming 128-bit loads): $(count xtb_static_reduce.o 'LDG.E.NA.128.CONSTANT')"
echo "xtb_static_f32.o  STG.E.EF.128 / STG.*128 (128-bit stores)         : $(count xtb_static_f32.o 'STG.E.*128')"
echo
echo "# hot loop of k_reduce_outer<copy_f32, sum> (unpredicated batch: 8 independent 128-bit loads, then the adds)"
cuobjdump -sass $L/xtb_static_reduce.o 2>/dev/null | awk '/Function : .*k_reduce_outerINS_10StaticEval.*Li2EEENS_9StaticAccILi64ELi5/{f=1} f' | grep "/\*[0-9a-f]\{4\}\*/" | sed 's@ */\* 0x[0-9a-f]* \*/@@' | grep -m1 -B2 -A22 "LDG.E.NA.128.CONSTANT" | cut -c1-100
