#include "scene.h"
#include <cstdio>
using namespace ptb;
int main(int argc, char** argv)
{
	// argv[1] = asset root, then scene json paths
	int ok = 0, bad = 0;
	for (int i = 2; i < argc; i++)
	{
		HostScene s;
		try { if (load_scene(argv[i], argv[1], s)) ok++; else bad++; } catch (const std::exception& e) { bad++; printf("exception %s on %s\n", e.what(), argv[i]); }
	}
	printf("ok %d bad %d\n", ok, bad);
	return 0;
}
