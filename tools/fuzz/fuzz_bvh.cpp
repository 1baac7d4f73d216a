#include "scene.h"
#include "bvh.h"
#include <cstdio>
using namespace ptb;
int main(int argc, char** argv)
{
	int ok = 0, bad = 0;
	for (int i = 2; i < argc; i++)
	{
		HostScene s;
		if (!load_scene(argv[i], argv[1], s)) { bad++; continue; }
		if (s.triangles.empty()) { ok++; continue; }
		for (int leaf : { 8, 3, 1 })
		{
			Bvh2 b; build_bvh2_sah(s.triangles, leaf, b, 0.8f);
			GpuBvh2 flat; flatten_bvh2(b, s.triangles, flat);
			if (leaf <= 3) { GpuBvh8 wide; build_bvh8(b, s.triangles, wide); }
		}
		ok++;
	}
	printf("ok %d bad %d\n", ok, bad);
}
