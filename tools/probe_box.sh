#!/bin/bash
# What the GPU box has: Fortran compilers, netCDF, host cores / NUMA, GPU topology.  Output -> gpurun_out/probe_box.txt
{
echo "== compilers"; for c in gfortran gfortran-11 gfortran-12 gfortran-13 gfortran-14 flang flang-new ifx ifort nvfortran pgfortran f95 f77; do printf "%s: " $c; command -v $c || echo absent; done
echo "== netcdf"; command -v nf-config nc-config ncdump || true; ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -i -E "netcdf|hdf5|gfortran" || echo "no netcdf/hdf5/libgfortran in /usr/lib/x86_64-linux-gnu"
find / -xdev \( -name "libgfortran*" -o -name "netcdf.mod" -o -name "libnetcdff*" \) 2>/dev/null | head
echo "== cpu"; nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core|Flags" | cut -c1-300
echo "== mem"; free -g | head -2
echo "== gpu"; nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm,clocks.max.mem --format=csv
nvidia-smi topo -m 2>/dev/null | head -20
echo "== gcc"; gcc --version | head -1; gcc -march=native -Q --help=target 2>/dev/null | grep -E "march|mavx512f|mfma " | head
} > gpurun_out/probe_box.txt 2>&1
