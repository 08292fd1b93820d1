#!/bin/sh
# One-command cross-check against the REAL reference for someone who has GHC (this image has none):
#   tools/ghc_crosscheck.sh /path/to/plonky2-verifier            (a checkout of bkomuves/plonky2-verifier)
# For every bundled fixture it points the reference's driver (src/testmain.hs reads ../json/<prefix>_{common,vkey,proof}.json)
# at the fixture, runs it with runghc, and compares with tests/golden/<name>.testmain.txt (what this repo's restatements say
# the reference prints).  A proof that ends in one of the reference's `error` sites is expected to print every line but the
# last on stdout and the message of the last line ("testmain: <message>") on stderr.
set -e
REF=${1:?usage: tools/ghc_crosscheck.sh /path/to/plonky2-verifier}
HERE=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$REF/json"
TMP=$(mktemp -d)
fail=0
sed 's/let prefix = "multi_lookup"/let prefix = "p2v"/' "$REF/src/testmain.hs" > "$REF/src/testmain_p2v.hs"
for proof in "$HERE"/tests/golden/*_proof.json; do
  name=$(basename "$proof" _proof.json)
  common=$name
  case $name in small6_bad*) common=small6 ;; real5_bad*) common=real5 ;; reallu6_bad*) common=reallu6 ;; esac
  cp "$HERE/tests/golden/${common}_common.json" "$REF/json/p2v_common.json"
  cp "$HERE/tests/golden/${name}_vkey.json" "$REF/json/p2v_vkey.json"
  cp "$proof" "$REF/json/p2v_proof.json"
  (cd "$REF/src" && runghc testmain_p2v.hs > "$TMP/out" 2> "$TMP/err") || true
  want="$HERE/tests/golden/$name.testmain.txt"
  if grep -q '^proof verification result = testmain: ' "$want"; then
    msg=$(sed -n 's/^proof verification result = testmain: //p' "$want")
    head -n -1 "$want" > "$TMP/want_head"
    grep -v '^proof verification result = ' "$TMP/out" > "$TMP/out_head" || true
    if cmp -s "$TMP/want_head" "$TMP/out_head" && grep -qF "$msg" "$TMP/err"; then echo "$name: identical (error site: $msg)"; else echo "$name: DIFFERS"; fail=1; fi
  else
    if cmp -s "$want" "$TMP/out"; then echo "$name: identical"; else echo "$name: DIFFERS"; fail=1; fi
  fi
done
rm -rf "$TMP" "$REF/src/testmain_p2v.hs" "$REF/json/p2v_common.json" "$REF/json/p2v_vkey.json" "$REF/json/p2v_proof.json"
exit $fail
