#!/bin/sh
# Installs the UNMODIFIED reference (KovenYu/uORF, pure Python) as oracle/_ref - the copy that travels to the GPU box.
# Run in the build container (needs /root/reference); __graft_entry__.build() does the same.  See oracle/ref_harness.py.
cd "$(dirname "$0")/.." && python -m oracle.ref_harness install "$@" && python -m oracle.ref_harness verify
