export OPMB200_LIB=$PWD/opm_simulators_b200/libopmb200_prof.so
timeout 100 python scripts/prof_tiles.py C3 1.0 dilu 4 2>&1 | head -4
DIMS=60x216x4 timeout 100 python scripts/prof_tiles.py C3 1.0 dilu 4 -804 2>&1 | head -4
DIMS=60x8x84 timeout 100 python scripts/prof_tiles.py C3 1.0 dilu 4 -804 2>&1 | head -4
