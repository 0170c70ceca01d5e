"""Does this box expose NVSwitch multicast (NVLS) to CUDA?  Prints the relevant device attributes."""
from cuda import cuda
err, = cuda.cuInit(0)
err, n = cuda.cuDeviceGetCount()
for i in range(n):
  err, dev = cuda.cuDeviceGet(i)
  vals = {}
  for name in ('CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED', 'CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED',
               'CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED', 'CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED'):
    attr = getattr(cuda.CUdevice_attribute, name, None)
    if attr is None:
      vals[name] = 'n/a'
      continue
    err, v = cuda.cuDeviceGetAttribute(attr, dev)
    vals[name.replace('CU_DEVICE_ATTRIBUTE_', '')] = (int(err), v)
  print(i, vals)
