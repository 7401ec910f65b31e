#!/bin/bash
# DEVELOPER TOOL (see shim_patch.py): run the Python side of a GPU test file on the CPU with the libgda calls replaced by
# torch expressions.  The file is copied to a scratch directory with "cuda" -> "cpu" substitutions; nothing here is
# collected by the repo's pytest runs.   bash tests/devtools/desk_check.sh tests/test_zz2_gpu_strurw.py [-k expr]
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
src=$1; shift
tmp=$(mktemp -d)
cp "$here/shim_patch.py" "$tmp/"
sed -e "s#^ROOT = .*#ROOT = \"$root\"#" "$root/tests/conftest.py" > "$tmp/conftest.py"
{ echo "import shim_patch  # noqa"; sed -e 's/"cuda:0"/"cpu"/g; s/\.cuda()/.cpu()/g; s/^pytestmark.*$//; s/is_cuda/is_cpu/g' "$root/$src"; } > "$tmp/test_desk.py"
cd "$tmp" && python -m pytest test_desk.py -x -q -p no:cacheprovider "$@"
