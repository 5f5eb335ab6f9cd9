/* oracle/stubs/hdf5.h -- TEST INFRASTRUCTURE ONLY: type names so raytrace.h parses; no HDF5 I/O is built. */
#ifndef ORACLE_STUB_HDF5_H
#define ORACLE_STUB_HDF5_H
typedef long hid_t; typedef int herr_t; typedef unsigned long long hsize_t;
#endif
