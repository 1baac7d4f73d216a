#include "scene.h"
#include <cstdio>
#include <cstring>
using namespace ptb;
int main(int argc, char** argv)
{
	// argv[1] = asset root, then scene json paths.  Every scene is loaded twice: OBJ text parsed in one pass, and cut
	// into 7 slices parsed by host threads; outcome, message and triangles must not depend on the slicing.
	int ok = 0, bad = 0, mismatch = 0;
	for (int i = 2; i < argc; i++)
	{
		HostScene s, s7;
		bool r = false, r7 = false;
		std::string e1, e7;
		try
		{
			set_loader_threads(1); r = load_scene(argv[i], argv[1], s); if (!r) e1 = last_error();
			set_loader_threads(7); r7 = load_scene(argv[i], argv[1], s7); if (!r7) e7 = last_error();
		}
		catch (const std::exception& e) { bad++; printf("exception %s on %s\n", e.what(), argv[i]); continue; }
		if (r) ok++; else bad++;
		const bool same = r == r7 && e1 == e7 && (!r || (s.triangles.size() == s7.triangles.size() && s.triangle_material == s7.triangle_material &&
			(s.triangles.empty() || memcmp(s.triangles.data(), s7.triangles.data(), s.triangles.size() * sizeof(Triangle)) == 0)));
		if (!same) { mismatch++; printf("exception sliced parse differs on %s: %d '%s' vs %d '%s'\n", argv[i], (int)r, e1.c_str(), (int)r7, e7.c_str()); }
	}
	printf("ok %d bad %d sliced-mismatch %d\n", ok, bad, mismatch);
	return 0;
}
