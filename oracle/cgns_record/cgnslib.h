/* TEST INFRASTRUCTURE — a RECORDING stand-in for the CGNS mid-level library (absent from this
 * image).  It declares only the handful of calls the reference's plot writers make
 * (DisjointBoxLayout.cpp:224-383, LevelData.cpp:65-196, LBLevel.H:96-188), with the argument
 * lists of the public CGNS API, and oracle/ref_lb_shim.cpp implements them by appending what
 * would have gone into the file to an in-memory record.  That pins WHICH cells, in WHICH order,
 * the reference writes (the valid region of every fab, ghosts stripped through
 * memrmin/memrmax), which is what the device-side plot gather has to reproduce.
 */
#ifndef SGC_CGNS_RECORD_H
#define SGC_CGNS_RECORD_H

typedef int cgsize_t;
enum ZoneType_t { ZoneTypeNull, ZoneTypeUserDefined, Structured, Unstructured };
enum GridLocation_t { GridLocationNull, GridLocationUserDefined, Vertex, CellCenter };
enum DataType_t { DataTypeNull, DataTypeUserDefined, Integer, RealSingle, RealDouble };
#define CG_MODE_READ 0
#define CG_MODE_WRITE 1

#ifdef __cplusplus
extern "C" {
#endif
int cg_open(const char* filename, int mode, int* fn);
int cg_close(int fn);
int cg_base_write(int fn, const char* basename, int cell_dim, int phys_dim, int* B);
int cg_zone_write(int fn, int B, const char* zonename, const cgsize_t* size, enum ZoneType_t type, int* Z);
int cg_coord_write(int fn, int B, int Z, enum DataType_t type, const char* coordname, const void* coord_ptr, int* C);
int cg_sol_write(int fn, int B, int Z, const char* solname, enum GridLocation_t location, int* S);
int cg_field_general_write(int fn, int B, int Z, int S, enum DataType_t type, const char* fieldname,
                           const cgsize_t* rmin, const cgsize_t* rmax, int m_numdim, const cgsize_t* m_dims,
                           const cgsize_t* m_rmin, const cgsize_t* m_rmax, const void* field_ptr, int* F);
void cg_error_print(void);
#ifdef __cplusplus
}
#endif
#endif
