#include "scene.h"
#include <cstdio>
#include <exception>
using namespace ptb;
int main(int argc, char** argv)
{
	HostScene s;
	try
	{
		bool ok = load_scene(argv[1], argv[2], s);
		printf("ok %d tris %zu %s\n", (int)ok, s.triangles.size(), ok ? "" : last_error().c_str());
	}
	catch (const std::exception& e) { printf("caught %s\n", e.what()); }
}
