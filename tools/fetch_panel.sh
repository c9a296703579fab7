#!/bin/sh
# data/human_sites_n10.fa.gz is the reference's shipped SNP panel (MIT-licensed DATA file
# /root/reference/data/human_sites_n10.fa, 96 287 sites / 1 270 317 k-mers), re-compressed so the
# benchmark and the full-panel parity tests can run on the GPU box where /root/reference is absent.
# ntsmCount reads it directly (gzopen), exactly like the reference would.
set -e
cd "$(dirname "$0")/.."
mkdir -p data
gzip -9 -n -c /root/reference/data/human_sites_n10.fa > data/human_sites_n10.fa.gz
ls -la data/human_sites_n10.fa.gz
