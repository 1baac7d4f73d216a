#include "scene.h"
#include <cstdio>
using namespace ptb;
int main(int argc, char** argv)
{
	int ok = 0, bad = 0;
	for (int i = 1; i < argc; i++) { Config c; try { if (load_config(argv[i], c)) ok++; else bad++; } catch (const std::exception& e) { bad++; printf("exception %s\n", e.what()); } }
	printf("ok %d bad %d\n", ok, bad);
}
