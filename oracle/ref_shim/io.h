/* Test-infrastructure shim for the MSVC <io.h> directory scan the reference uses at
 * Core/scene_parser.cpp:15-27 (_findfirst/_findnext/_findclose over "dir\\*.json").
 * Implemented over POSIX glob(); '\\' separators are mapped to '/'. Not product code. */
#pragma once
#include <glob.h>
#include <string.h>
#include <stdint.h>
#include <string>

struct _finddata_t { char name[260]; };

struct ptb_find_state { glob_t g; size_t next; };

static inline void ptb_fill_name(_finddata_t* info, const char* full)
{
	const char* base = strrchr(full, '/');
	base = base ? base + 1 : full;
	strncpy(info->name, base, sizeof(info->name) - 1);
	info->name[sizeof(info->name) - 1] = 0;
}

static inline intptr_t _findfirst(const char* pattern, _finddata_t* info)
{
	std::string p(pattern);
	for (auto& c : p) if (c == '\\') c = '/';
	ptb_find_state* st = new ptb_find_state();
	st->next = 0;
	info->name[0] = '.'; info->name[1] = 0;
	if (glob(p.c_str(), 0, nullptr, &st->g) != 0 || st->g.gl_pathc == 0) { return (intptr_t)st; }
	ptb_fill_name(info, st->g.gl_pathv[0]);
	st->next = 1;
	return (intptr_t)st;
}

static inline int _findnext(intptr_t h, _finddata_t* info)
{
	ptb_find_state* st = (ptb_find_state*)h;
	if (st->next >= st->g.gl_pathc) return -1;
	ptb_fill_name(info, st->g.gl_pathv[st->next++]);
	return 0;
}

static inline int _findclose(intptr_t h)
{
	ptb_find_state* st = (ptb_find_state*)h;
	globfree(&st->g);
	delete st;
	return 0;
}
