"""TEST INFRASTRUCTURE — a numpy model of the two-steps-per-pass sweep's ALGORITHM (the data flow of
structuredgridcalc-boxframework_b200/csrc/sgc_stream2.cu, none of its CUDA): per box, tiles of level t
assembled from the neighbours' VALID cells (or, where a face has no motion item, from the never-exchanged
ghost cells of the step-t / step-t+1 levels), level t+1 on the valid cells plus the four one-cell flanks,
level t+2 on the valid cells.  tests/test_two_step_model.py checks it against two { exchange ; sweep }
steps of the oracle (LevelData.H:455-566 + the sweep loop): the claim behind the kernel — every t+1 value
a box needs can be recomputed from valid cells and static ghosts, bit for bit — holds on the CPU too.

Arithmetic: the reference loop's association without contraction (oracle contract = 0).

What the model leaves out on purpose, because it changes no value: the ORDER in which the kernel visits planes
(it marches down pillars of z-stacked boxes and reads the planes two boxes share once) and its form
fma(-2, u, u[+e]) of (u[+e] - 2u), which is exact (tests/test_two_step_model.py::test_fma_form_of_the_update_is_exact).
"""
import numpy as np


def upd(c, xm, xp, ym, yp, zm, zp, f):
    twoc = 2.0 * c
    tx = (xp - twoc) + xm
    ty = (yp - twoc) + ym
    tz = (zp - twoc) + zm
    return c + f * ((tx + ty) + tz)


def face_neighbours(items):
    """(box, (dx,dy,dz)) -> the box whose valid cells fill that face's ghost cells (face items only)"""
    nbr = {}
    for it in items:
        d = (int(it[14]), int(it[15]), int(it[16]))
        if abs(d[0]) + abs(d[1]) + abs(d[2]) == 1:
            nbr[(int(it[0]), d)] = int(it[1])
    return nbr


def split(flat, boxes, ng=1):
    """flat level (fabs concatenated, reference order) -> list of [k, j, i] arrays (views)"""
    out, off = [], 0
    for b in boxes:
        n = [int(b[3 + d] - b[d] + 1 + 2 * ng) for d in range(3)]
        out.append(flat[off:off + n[0] * n[1] * n[2]].reshape(n[2], n[1], n[0]))
        off += n[0] * n[1] * n[2]
    return out


def two_step_pass(cur, gA, gB, nbr, f):
    """cur: fabs holding level t in their valid cells; gA / gB: fabs whose ghost cells are read as level t /
    level t+1 where a face has no neighbour.  Returns the valid cells of level t+2, one [vz, V, V] array per box.
    Fabs are [k, j, i] with one ghost layer; boxes must be square in x-y (V x V x vz)."""
    nbox = len(cur)
    res = []
    for b in range(nbox):
        vz, V = cur[b].shape[0] - 2, cur[b].shape[1] - 2
        assert cur[b].shape[2] - 2 == V
        xm_b, xp_b = nbr.get((b, (-1, 0, 0))), nbr.get((b, (1, 0, 0)))
        ym_b, yp_b = nbr.get((b, (0, -1, 0))), nbr.get((b, (0, 1, 0)))
        zm_b, zp_b = nbr.get((b, (0, 0, -1))), nbr.get((b, (0, 0, 1)))

        def tile(z):
            """[(y+2), (x+2)] for y, x in -2..V+1; NaN where the kernel's slot holds nothing meaningful"""
            T = np.full((V + 4, V + 4), np.nan)
            zmiss = False
            if 0 <= z < vz:
                S, fp, lev = b, z + 1, cur
            else:
                nb = zm_b if z < 0 else zp_b
                if nb is None:                  # no z neighbour: the box's own ghost plane, of the step-t ghosts
                    zmiss = True                # for the inner halo plane, of the step-t+1 ghosts for the outer one
                    S, fp = b, (0 if z < 0 else vz + 1)
                    lev = gA if z in (-1, vz) else gB
                else:
                    S, fp, lev = nb, (z + vz if z < 0 else z - vz) + 1, cur
            P = lev[S][fp]
            T[2:V + 2, 2:V + 2] = P[1:V + 1, 1:V + 1]
            for dy, rows_t, rows_n, grow in ((-1, (0, 1), (V - 1, V), 0), (1, (V + 2, V + 3), (1, 2), V + 1)):
                N = nbr.get((S, (0, dy, 0)))
                if N is not None:
                    T[rows_t[0], 2:V + 2] = lev[N][fp][rows_n[0], 1:V + 1]
                    T[rows_t[1], 2:V + 2] = lev[N][fp][rows_n[1], 1:V + 1]
                elif not zmiss:                 # the step-t ghost row inside, the step-t+1 ghost row in the outer slot
                    inner, outer = (1, 0) if dy < 0 else (V + 2, V + 3)
                    T[inner, 2:V + 2] = gA[S][fp][grow, 1:V + 1]
                    T[outer, 2:V + 2] = gB[S][fp][grow, 1:V + 1]
            for dx, cols_t, cols_n, gcol in ((-1, (0, 1), (V - 1, V), 0), (1, (V + 2, V + 3), (1, 2), V + 1)):
                N = nbr.get((S, (dx, 0, 0)))
                if N is not None:
                    if zmiss:                   # ghost plane: only the column next to the box is needed
                        T[2:V + 2, 1 if dx < 0 else V + 2] = lev[N][fp][1:V + 1, V if dx < 0 else 1]
                    else:                       # (the edge strips of the x neighbour)
                        T[2:V + 2, cols_t[0]] = lev[N][fp][1:V + 1, cols_n[0]]
                        T[2:V + 2, cols_t[1]] = lev[N][fp][1:V + 1, cols_n[1]]
                elif not zmiss:
                    inner, outer = (1, 0) if dx < 0 else (V + 2, V + 3)
                    T[2:V + 2, inner] = gA[S][fp][1:V + 1, gcol]
                    T[2:V + 2, outer] = gB[S][fp][1:V + 1, gcol]
            if not zmiss:
                for dx in (-1, 1):
                    for dy in (-1, 1):
                        X, Y = nbr.get((S, (dx, 0, 0))), nbr.get((S, (0, dy, 0)))
                        ty_, tx_ = (1 if dy < 0 else V + 2), (1 if dx < 0 else V + 2)
                        if X is not None and Y is not None:     # the diagonal box, over two faces
                            D = nbr[(X, (0, dy, 0))]
                            T[ty_, tx_] = cur[D][fp][V if dy < 0 else 1, V if dx < 0 else 1]
                        elif X is not None:                     # y face missing: the x neighbour's ghost row
                            T[ty_, tx_] = gA[X][fp][0 if dy < 0 else V + 1, V if dx < 0 else 1]
                        elif Y is not None:                     # x face missing: the y neighbour's ghost column
                            T[ty_, tx_] = gA[Y][fp][V if dy < 0 else 1, 0 if dx < 0 else V + 1]
            return T

        tiles = {z: tile(z) for z in range(-2, vz + 2)}
        c_ = slice(2, V + 2)
        # stage 1: level t+1 on [(y+1), (x+1)] for y, x in -1..V (valid cells + flanks, no corners)
        t1 = {}
        for z in range(-1, vz + 1):
            U = np.full((V + 2, V + 2), np.nan)
            T, Tm, Tp = tiles[z], tiles[z - 1], tiles[z + 1]
            if z == -1 and zm_b is None:
                U[1:V + 1, 1:V + 1] = Tm[c_, c_]            # the step-t+1 ghost plane came in the outer tile
            elif z == vz and zp_b is None:
                U[1:V + 1, 1:V + 1] = Tp[c_, c_]
            else:
                U[1:V + 1, 1:V + 1] = upd(T[c_, c_], T[c_, 1:V + 1], T[c_, 3:V + 3], T[1:V + 1, c_], T[3:V + 3, c_],
                                          Tm[c_, c_], Tp[c_, c_], f)
            if 0 <= z < vz:                                 # flanks: computed, or the box's step-t+1 ghost cells
                for col, tcol, miss, outer in ((0, 1, xm_b is None, 0), (V + 1, V + 2, xp_b is None, V + 3)):
                    U[1:V + 1, col] = T[c_, outer] if miss else upd(T[c_, tcol], T[c_, tcol - 1], T[c_, tcol + 1],
                                                                    T[1:V + 1, tcol], T[3:V + 3, tcol], Tm[c_, tcol],
                                                                    Tp[c_, tcol], f)
                for row, trow, miss, outer in ((0, 1, ym_b is None, 0), (V + 1, V + 2, yp_b is None, V + 3)):
                    U[row, 1:V + 1] = T[outer, c_] if miss else upd(T[trow, c_], T[trow, 1:V + 1], T[trow, 3:V + 3],
                                                                    T[trow - 1, c_], T[trow + 1, c_], Tm[trow, c_],
                                                                    Tp[trow, c_], f)
            t1[z] = U
        # stage 2: level t+2 on the valid cells
        out = np.empty((vz, V, V))
        v_ = slice(1, V + 1)
        for z in range(vz):
            U = t1[z]
            out[z] = upd(U[v_, v_], U[v_, 0:V], U[v_, 2:V + 2], U[0:V, v_], U[2:V + 2, v_], t1[z - 1][v_, v_],
                         t1[z + 1][v_, v_], f)
        assert not np.isnan(out).any(), "the model read a slot nothing was put into"
        res.append(out)
    return res
