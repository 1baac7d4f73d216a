// Minimal JSON reader for the reference's config and scene files (flat
// objects / arrays whose scalars are all strings).  Full JSON value grammar is accepted so that
// hand-edited files with numbers or booleans produce a clear type error instead of a crash.
#pragma once
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace ptb
{

struct JValue
{
	enum Type { Null, Bool, Number, String, Array, Object } type = Null;
	bool b = false;
	double num = 0.0;
	std::string str;
	std::vector<JValue> arr;
	std::vector<std::pair<std::string, JValue>> obj;

	bool is_null() const { return type == Null; }
	bool is_array() const { return type == Array; }
	bool is_object() const { return type == Object; }
	bool is_string() const { return type == String; }
	// nlohmann's operator[] semantics on a missing key: a null value.  Last duplicate wins.
	const JValue& operator[](const char* key) const
	{
		static const JValue null_value;
		if (type != Object) return null_value;
		const JValue* found = &null_value;
		for (auto& kv : obj) if (kv.first == key) found = &kv.second;
		return *found;
	}
};

class JParser
{
public:
	explicit JParser(const std::string& text) : s(text), i(0) {}
	bool parse(JValue& out, std::string& err)
	{
		skip_bom();
		if (!value(out)) { err = error.empty() ? "syntax error" : error; err += " at byte " + std::to_string(i); return false; }
		ws();
		if (i != s.size()) { err = "trailing characters at byte " + std::to_string(i); return false; }
		return true;
	}

private:
	const std::string& s;
	size_t i;
	std::string error;

	void skip_bom() { if (s.size() >= 3 && (unsigned char)s[0] == 0xEF && (unsigned char)s[1] == 0xBB && (unsigned char)s[2] == 0xBF) i = 3; }
	void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++; }
	bool fail(const char* m) { if (error.empty()) error = m; return false; }

	bool value(JValue& v)
	{
		ws();
		if (i >= s.size()) return fail("unexpected end of input");
		char c = s[i];
		if (c == '{') return object(v);
		if (c == '[') return array(v);
		if (c == '"') { v.type = JValue::String; return string(v.str); }
		if (!s.compare(i, 4, "true")) { v.type = JValue::Bool; v.b = true; i += 4; return true; }
		if (!s.compare(i, 5, "false")) { v.type = JValue::Bool; v.b = false; i += 5; return true; }
		if (!s.compare(i, 4, "null")) { v.type = JValue::Null; i += 4; return true; }
		if (c == '-' || (c >= '0' && c <= '9'))
		{
			char* end = nullptr;
			v.num = strtod(s.c_str() + i, &end);
			if (end == s.c_str() + i) return fail("bad number");
			i = end - s.c_str();
			v.type = JValue::Number;
			return true;
		}
		return fail("unexpected character");
	}

	bool string(std::string& out)
	{
		i++; // opening quote
		out.clear();
		while (i < s.size())
		{
			char c = s[i++];
			if (c == '"') return true;
			if (c != '\\') { out.push_back(c); continue; }
			if (i >= s.size()) break;
			char e = s[i++];
			switch (e)
			{
			case '"': out.push_back('"'); break;
			case '\\': out.push_back('\\'); break;
			case '/': out.push_back('/'); break;
			case 'b': out.push_back('\b'); break;
			case 'f': out.push_back('\f'); break;
			case 'n': out.push_back('\n'); break;
			case 'r': out.push_back('\r'); break;
			case 't': out.push_back('\t'); break;
			case 'u':
			{
				if (i + 4 > s.size()) return fail("bad \\u escape");
				unsigned cp = (unsigned)strtoul(s.substr(i, 4).c_str(), nullptr, 16);
				i += 4;
				if (cp < 0x80) out.push_back((char)cp);
				else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
				else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
				break;
			}
			default: return fail("bad escape");
			}
		}
		return fail("unterminated string");
	}

	bool array(JValue& v)
	{
		v.type = JValue::Array;
		i++;
		ws();
		if (i < s.size() && s[i] == ']') { i++; return true; }
		while (true)
		{
			v.arr.emplace_back();
			if (!value(v.arr.back())) return false;
			ws();
			if (i < s.size() && s[i] == ',') { i++; continue; }
			if (i < s.size() && s[i] == ']') { i++; return true; }
			return fail("expected ',' or ']'");
		}
	}

	bool object(JValue& v)
	{
		v.type = JValue::Object;
		i++;
		ws();
		if (i < s.size() && s[i] == '}') { i++; return true; }
		while (true)
		{
			ws();
			if (i >= s.size() || s[i] != '"') return fail("expected object key");
			std::string key;
			if (!string(key)) return false;
			ws();
			if (i >= s.size() || s[i] != ':') return fail("expected ':'");
			i++;
			v.obj.emplace_back(key, JValue());
			if (!value(v.obj.back().second)) return false;
			ws();
			if (i < s.size() && s[i] == ',') { i++; continue; }
			if (i < s.size() && s[i] == '}') { i++; return true; }
			return fail("expected ',' or '}'");
		}
	}
};

} // namespace ptb
