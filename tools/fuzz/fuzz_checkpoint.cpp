#include "image_out.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
using namespace ptb;
int main()
{
	CheckpointHeader h; memset(&h, 0, sizeof(h)); h.width = 13; h.height = 7; h.pass_counter = 5; h.max_depth = 8;
	std::vector<float> sum(13 * 7 * 3, 0.25f);
	std::string err;
	if (!write_checkpoint("/tmp/ptb_fuzz_c.bin", h, sum.data(), err)) { printf("write failed %s\n", err.c_str()); return 1; }
	std::ifstream f("/tmp/ptb_fuzz_c.bin", std::ios::binary); std::vector<char> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	std::mt19937 rng(3);
	int ok = 0, bad = 0;
	for (int it = 0; it < 20000; it++)
	{
		std::vector<char> a = data;
		int k = 1 + rng() % 4;
		for (int j = 0; j < k; j++)
		{
			int m = rng() % 3; size_t p = rng() % a.size();
			if (m == 0) a[p] = (char)rng();
			else if (m == 1) a.erase(a.begin() + p, a.begin() + std::min(a.size(), p + 1 + rng() % 40));
			else a.insert(a.begin() + p, (size_t)(1 + rng() % 8), (char)rng());
			if (a.empty()) a.push_back(0);
		}
		std::ofstream o("/tmp/ptb_fuzz_m.bin", std::ios::binary); o.write(a.data(), a.size()); o.close();
		CheckpointHeader h2; std::vector<float> s2;
		if (read_checkpoint("/tmp/ptb_fuzz_m.bin", h2, s2, err)) ok++; else bad++;
	}
	printf("ok %d bad %d\n", ok, bad);
}
