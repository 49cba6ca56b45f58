#!/bin/sh
# Recipe for oracle/_ref/: the reference's OWN files of the hot path, taken from where they lie under
# /root/reference at build time (this container only; the GPU box uses the copies that travel with the snapshot).
# oracle/_ref/ is git-ignored: nothing of the reference enters the history.  bench.py's reference arm and
# cpu_baseline leg import these two modules when they are present (kind "reference") and fall back to
# oracle/scipy_chain.py (kind "port") otherwise.  Both files are pure Python on numpy + scipy.ndimage.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MAZE_REFERENCE:-/root/reference}"
[ -f "$REF/maze_ipp/isotropic.py" ] || { echo "make_ref.sh: $REF/maze_ipp/isotropic.py not found" >&2; exit 1; }
mkdir -p "$HERE/_ref"
install -m 0644 "$REF/maze_ipp/isotropic.py" "$HERE/_ref/isotropic.py"
install -m 0644 "$REF/maze_ipp/merge_labels.py" "$HERE/_ref/merge_labels.py"
echo "oracle/_ref: isotropic.py merge_labels.py (from $REF)"
