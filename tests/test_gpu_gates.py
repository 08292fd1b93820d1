"""GPU: every gate program, switched on alone through its selector, vanishes on an honest witness row and fires
on a broken one — through p2v_constraints on a real proof blob (the mid5 fixture with its openings replaced),
bit-exact against the oracle's combined value."""
import json

import numpy as np
import pytest

import fixtures
import test_honest_rows as hr

pytestmark = pytest.mark.gpu
P = hr.P


def test_each_gate_alone_on_honest_and_broken_rows(p2v, ctx, orc):
    shape, lay, vkey, blob = fixtures.load("mid5")
    common = json.loads(fixtures.read("mid5", "common"))
    cir = p2v.Circuit(ctx, shape, vkey)
    pis = blob[lay.off_public_inputs: lay.off_public_inputs + shape.num_public_inputs]
    pih = [int(x) for x in ctx.sponge(pis.reshape(-1, 1).copy())[:, 0]]
    blobs, expect_zero, names = [], [], []
    for k, text in enumerate(common["gates"]):
        g = hr.pyref.parse_gate(text)
        if g[0] == "NoopGate":
            continue
        consts = [hr.rnd_ext(), hr.rnd_ext()]
        w = hr.poseidon_gate_row(k % 2) if g[0] == "PoseidonGate" else hr.honest_row(g, consts)
        if g[0] == "PublicInputGate":
            for i in range(4):
                w[i] = hr.E(pih[i])
        for broken in (False, True):
            b = blob.copy()
            wires = np.array([e.pair() for e in w], dtype=np.uint64).reshape(-1)
            if broken:
                wires[0] = (int(wires[0]) + 1) % P
                if g[0] == "PoseidonGate":
                    wires[2 * 70] = (int(wires[2 * 70]) + 1) % P
            b[lay.off_open_wires: lay.off_open_wires + 2 * shape.num_wires] = wires
            # selector columns: this gate's group holds its index, the others UNUSED (Gate/Selector.hs:83-89)
            for grp in range(shape.num_groups):
                val = k if grp == shape.gates[k].group else 0xFFFFFFFF
                b[lay.off_open_constants + 2 * grp] = val
                b[lay.off_open_constants + 2 * grp + 1] = 0
            for i, c in enumerate(consts):
                b[lay.off_open_constants + 2 * (shape.num_groups + i): lay.off_open_constants + 2 * (shape.num_groups + i) + 2] = c.pair()
            blobs.append(b)
            expect_zero.append(not broken)
            names.append(g[0] + ("/broken" if broken else ""))
    blobs = np.stack(blobs)
    comb, mask = cir.evalCombinedPlonkConstraints(blobs)
    want = orc.verify_batch(shape, vkey, blobs, threads=4, fast=True)
    assert np.array_equal(comb, want["combined"])
    assert np.array_equal(mask, want["eqmask"])
    full = (1 << shape.num_challenges) - 1
    for i, nm in enumerate(names):
        is_zero = not comb[:, i].any()
        assert is_zero == expect_zero[i], nm
        assert (mask[i] == full) == expect_zero[i], nm
    assert len(names) == 26
