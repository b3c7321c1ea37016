#!/bin/bash
# Build the library of another revision next to the current one: tools/ab_build.sh <git-rev> <tag>
#   -> .ab/lib_<tag>.so (git-ignored, travels to the GPU box); use it with DM_LIB_PATH=.ab/lib_<tag>.so
set -e
cd "$(dirname "$0")/.."
rev=$1; tag=$2
mkdir -p .ab
rm -rf .ab/wt && git worktree prune
git worktree add -f .ab/wt "$rev" > /dev/null 2>&1
(cd .ab/wt && python -m deepmerge_b200.build > /dev/null)
cp .ab/wt/deepmerge_b200/libdeepmerge_b200.so ".ab/lib_${tag}.so"
git worktree remove --force .ab/wt
echo "built .ab/lib_${tag}.so from $rev"
