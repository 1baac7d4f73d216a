#include "scene.h"
#include <cstdio>
using namespace ptb;
int main(int argc, char** argv)
{
	int ok = 0, bad = 0;
	for (int i = 1; i < argc; i++) { Texture t; if (load_image_rgba8(argv[i], t)) ok++; else bad++; }
	printf("ok %d bad %d\n", ok, bad);
	return 0;
}
