#!/bin/bash
# Probe the GPU box for any OpenMM install (the denominator of the north_star's 10x target is the
# plugin's stock CUDA platform, which needs OpenMM + its CUDA plugin).  Output goes to gpurun_out/.
out=gpurun_out/probe_openmm.log
mkdir -p gpurun_out
{
  echo "== date"; date -u
  echo "== nvidia-smi"; nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
  echo "== python -c 'import openmm'"; python -c "import openmm; print(openmm.__version__, openmm.__file__)" 2>&1 | tail -1
  echo "== python -c 'import simtk.openmm'"; python -c "import simtk.openmm" 2>&1 | tail -1
  echo "== python -c 'import nonbondedslicing'"; python -c "import nonbondedslicing" 2>&1 | tail -1
  echo "== pip list | grep -i openmm"; python -m pip list 2>/dev/null | grep -i -E "openmm|nonbonded" || echo "(none)"
  echo "== ls baseline/_ref"; ls -la baseline/_ref 2>&1 | head
  echo "== which conda mamba micromamba swig"; which conda mamba micromamba swig 2>&1
  echo "== find / -name 'libOpenMM*'"; find / -xdev \( -name 'libOpenMM*' -o -name 'OpenMM.h' -o -name 'openmm*.whl' -o -name 'openmm*.tar*' -o -name 'openmm*.conda' \) 2>/dev/null | head -20; echo "(end of find)"
  echo "== /opt/wheelhouse openmm"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "openmm|nonbonded" || echo "(none in /opt/wheelhouse)"
  echo "== conda dirs"; ls -d /opt/conda /root/miniconda3 /root/anaconda3 /usr/local/openmm /opt/openmm 2>&1
  echo "== network"; timeout 5 python - <<'PY' 2>&1 | tail -1
import socket
try:
    socket.create_connection(("pypi.org", 443), timeout=3); print("network: reachable")
except Exception as e:
    print("network: unreachable:", e)
PY
} > $out 2>&1
cat $out
