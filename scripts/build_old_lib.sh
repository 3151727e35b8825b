#!/bin/bash
# Builds the library of an earlier commit next to the tree's own one, for same-box A/B runs:
#   scripts/build_old_lib.sh <commit> <name>   ->  ark_analysis_b200/_lib/libpixie_b200_<name>.so
# A/B scripts select it with PIXIE_LIB_PATH (ark_analysis_b200/_native.py); the file is git-ignored
# and travels to the GPU box with the tree.  Every A/B in profiles/r02_notes.md that says "against
# the previous build" was made this way -- a switch inside the new binary hides what the new code
# costs the paths that do not use it (section 12 there).
set -e
commit=${1:?commit}; name=${2:?name}
root=$(git rev-parse --show-toplevel)
wt=$(mktemp -d /tmp/pixie_wt.XXXXXX)
git -C "$root" worktree add -q "$wt" "$commit"
make -C "$wt/ark_analysis_b200/csrc" -j8 > /dev/null
cp "$wt/ark_analysis_b200/_lib/libpixie_b200.so" "$root/ark_analysis_b200/_lib/libpixie_b200_$name.so"
git -C "$root" worktree remove --force "$wt"
ls -la "$root/ark_analysis_b200/_lib/libpixie_b200_$name.so"
