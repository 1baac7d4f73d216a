/* Fixture-generator helper (NOT product code): writes the OBJ text of pathtracerwithcuda_b200/procedural.py's write_obj() with C stdio
 * instead of numpy.savetxt — byte-identical output ("%.6f" of the same doubles, "%d//%d" index triples), ~15x faster, which matters for the
 * 5 M-triangle workload c5 (374 MB of text) that bench.py generates on the box.  Built by pathtracerwithcuda_b200/build.py with gcc into
 * tools/fastobj/libfastobj.so; procedural.py falls back to numpy.savetxt when the library is absent (tests/test_frontend.py compares both). */
#include <stdio.h>
#include <stdint.h>

/* appends to `path`: n_v "v" lines, n_vt "vt" lines (uvs may be NULL), n_v "vn" lines */
int fastobj_vertices(const char* path, const double* verts, const double* uvs, const double* normals, int64_t n_v, int64_t n_vt)
{
	FILE* f = fopen(path, "a");
	if (!f) return 1;
	static char buf[1 << 20];
	setvbuf(f, buf, _IOFBF, sizeof(buf));
	for (int64_t i = 0; i < n_v; i++) fprintf(f, "v %.6f %.6f %.6f\n", verts[i * 3], verts[i * 3 + 1], verts[i * 3 + 2]);
	if (uvs) for (int64_t i = 0; i < n_vt; i++) fprintf(f, "vt %.6f %.6f\n", uvs[i * 2], uvs[i * 2 + 1]);
	for (int64_t i = 0; i < n_v; i++) fprintf(f, "vn %.6f %.6f %.6f\n", normals[i * 3], normals[i * 3 + 1], normals[i * 3 + 2]);
	return fclose(f) != 0;
}

/* appends "g <name>" and n faces (1-based indices in `faces`, 3 per face); with_uv: "f a/a/a b/b/b c/c/c" else "f a//a b//b c//c" */
int fastobj_group(const char* path, const char* name, const int64_t* faces, int64_t n, int with_uv)
{
	FILE* f = fopen(path, "a");
	if (!f) return 1;
	static char buf[1 << 20];
	setvbuf(f, buf, _IOFBF, sizeof(buf));
	fprintf(f, "g %s\n", name);
	for (int64_t i = 0; i < n; i++)
	{
		const long long a = faces[i * 3], b = faces[i * 3 + 1], c = faces[i * 3 + 2];
		if (with_uv) fprintf(f, "f %lld/%lld/%lld %lld/%lld/%lld %lld/%lld/%lld\n", a, a, a, b, b, b, c, c, c);
		else fprintf(f, "f %lld//%lld %lld//%lld %lld//%lld\n", a, a, b, b, c, c);
	}
	return fclose(f) != 0;
}
