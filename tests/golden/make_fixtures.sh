#!/bin/sh
# Regenerates the bundled JSON fixtures with the trivial-circuit mini-prover (oracle/prover.cpp).
# The reference bundles none (its json/ directory is git-ignored), so these stand in for
# "the repo's bundled Plonky2 JSON proof" of BASELINE.json configs[0].  Deterministic (seeded).
set -e
cd "$(dirname "$0")/../.."
make -C oracle -s p2v_prover
P=oracle/p2v_prover
G=tests/golden
$P s12     $G/s12     --seed 1
$P mid5    $G/mid5    --seed 2
$P small6  $G/small6  --seed 3
$P fixed4  $G/fixed4  --seed 4
$P lookup6 $G/lookup6 --seed 5
$P arity5  $G/arity5  --seed 7   # Fixed [5,1]: a 32-point coset per query
# real circuit: ACTIVE gates of all 14 standard kinds on honest witnesses, copy constraints (non-identity sigma,
# real grand product Z + partial products), real quotient polynomial
$P real5   $G/real5   --seed 6
$P real12  $G/real12  --seed 12  # the standard recursion configuration (2^12 rows, 28 queries, arity 16) with real rows
$P reallu6 $G/reallu6 --seed 8   # real circuit with an honest lookup argument (Lookup / LookupTable rows, RE + partial sums)
# rejecting proofs that need a prover-side change (re-grinding after the change):
$P small6 $G/small6_badfinal  --seed 3 --bad-final        # -> FALSE_FINAL
$P small6 $G/small6_badlayer0 --seed 3 --corrupt-layer 0  # -> ERR_STEP_EVAL(step 0)
$P small6 $G/small6_badlayer1 --seed 3 --corrupt-layer 1  # -> ERR_STEP_EVAL(step 1)
$P real5  $G/real5_badwitness  --seed 6 --bad-witness 12  # PoseidonGate row left unsatisfied -> FALSE_EQS
$P real5  $G/real5_badcopy     --seed 6 --bad-copy        # one wired cell differs from its cycle -> FALSE_EQS
$P reallu6 $G/reallu6_badlookup --seed 8 --bad-lookup      # one looked-up pair is not in the table -> FALSE_EQS
rm -f $G/real5_badwitness_common.json $G/real5_badcopy_common.json $G/reallu6_badlookup_common.json
# the variants share small6's circuit: keep one copy of common/vkey where identical
for v in badfinal badlayer0 badlayer1; do rm -f $G/small6_${v}_common.json; done
