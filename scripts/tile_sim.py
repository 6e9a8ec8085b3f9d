"""List-scheduling model of the tile walkers on a box grid: tiles of TJ x TK lines, in-order tickets, P persistent CTAs.
A tile starts when its CTA is free (+ start-up U) and its upstream tiles are 8 / 4 steps (+ one hop H) ahead; all steps
cost c cycles.  python scripts/tile_sim.py nx ny nz [c] [H] [U] [P] [TJ] [TK]"""
import heapq, sys
nx, ny, nz = (int(v) for v in sys.argv[1:4])
c = float(sys.argv[4]) if len(sys.argv) > 4 else 564
H = float(sys.argv[5]) if len(sys.argv) > 5 else 1700
U = float(sys.argv[6]) if len(sys.argv) > 6 else 4000
P = int(sys.argv[7]) if len(sys.argv) > 7 else 148
TJ = int(sys.argv[8]) if len(sys.argv) > 8 else 8
TK = int(sys.argv[9]) if len(sys.argv) > 9 else 4
ntj, ntk = -(-ny // TJ), -(-nz // TK)
tiles = []
for tk in range(ntk):
    for tj in range(ntj):
        lj, lk = min(TJ, ny - tj * TJ), min(TK, nz - tk * TK)
        tiles.append((TJ * tj + TK * tk, tj, tk, nx + lj + lk - 2))
tiles.sort()
free = [0.0] * P
heapq.heapify(free)
S, E = {}, {}
busy = 0.0
for w, tj, tk, ns in tiles:
    f = heapq.heappop(free) + U
    s = f
    if tj > 0:
        s = max(s, S[(tj - 1, tk)] + TJ * c + H)
    if tk > 0:
        s = max(s, S[(tj, tk - 1)] + TK * c + H)
    S[(tj, tk)] = s
    e = s + ns * c
    E[(tj, tk)] = e
    busy += ns * c
    heapq.heappush(free, e)
T = max(E.values())
cp = (nx + ny + nz - 2) * c + (ntj - 1 + ntk - 1) * H
print(f"tiles {len(tiles)} T = {T:.0f} cycles = {T/1965:.1f} us; critical path {cp:.0f} = {cp/1965:.1f} us; work/P {busy/P:.0f} = {busy/P/1965:.1f} us; utilisation {busy/P/T:.2f}")
