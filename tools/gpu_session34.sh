#!/bin/bash
O=gpurun_out/s34; mkdir -p $O
for c in 128 512 1024; do CHUNK=$c timeout 600 python tools/stream_probe.py 4096 > $O/streaming_windows_$c.md 2> $O/stream.err; cat $O/streaming_windows_$c.md; tail -3 $O/stream.err; done
