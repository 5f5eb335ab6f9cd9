/* oracle/stubs/hdf5_hl.h -- TEST INFRASTRUCTURE ONLY: type names so raytrace.h parses; no HDF5 I/O is built. */
#ifndef ORACLE_STUB_HDF5_HL_H
#define ORACLE_STUB_HDF5_HL_H
typedef long hid_t; typedef int herr_t; typedef unsigned long long hsize_t;
#endif
