// value_replay.cpp -- test driver for the Value-level ring helpers (Triple::sum_triple / subtract_triple /
// sum_nb_triple, imputation/include/sum_sub.h:10-14).  Like replay_host.cpp it is linked twice: with value_glue.cpp
// (this repo) and with the reference's own imputation/triple/{sum,sub,sum_nb}.cpp (oracle/_ref) -- one driver, two
// implementations behind the same three functions.  Values cross the C boundary as JSON: object = STRUCT (field
// order kept), array = LIST, number = scalar.
#include <duckdb.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace Triple {
duckdb::Value subtract_triple(duckdb::Value &triple_1, duckdb::Value &triple_2);
duckdb::Value sum_triple(const duckdb::Value &triple_1, const duckdb::Value &triple_2);
duckdb::Value sum_nb_triple(const duckdb::Value &triple_1, const duckdb::Value &triple_2);
}  // namespace Triple

namespace {
thread_local std::string g_value_error;

void Ws(const char *&p) {
  while (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') p++;
}

duckdb::Value Parse(const char *&p) {
  Ws(p);
  if (*p == '{') {
    p++;
    duckdb::child_list_t<duckdb::Value> fields;
    Ws(p);
    while (*p != '}') {
      if (*p != '"') throw duckdb::InvalidInputException("value json: expected a field name");
      const char *e = strchr(p + 1, '"');
      if (!e) throw duckdb::InvalidInputException("value json: unterminated field name");
      std::string name(p + 1, e);
      p = e + 1;
      Ws(p);
      if (*p++ != ':') throw duckdb::InvalidInputException("value json: expected ':'");
      duckdb::Value child = Parse(p);
      fields.emplace_back(name, std::move(child));
      Ws(p);
      if (*p == ',') p++;
      Ws(p);
    }
    p++;
    return duckdb::Value::STRUCT(std::move(fields));
  }
  if (*p == '[') {
    p++;
    duckdb::vector<duckdb::Value> items;
    Ws(p);
    while (*p != ']') {
      items.push_back(Parse(p));
      Ws(p);
      if (*p == ',') p++;
      Ws(p);
    }
    p++;
    return duckdb::Value::LIST(duckdb::LogicalType(), std::move(items));
  }
  char *end = nullptr;
  const double d = strtod(p, &end);
  if (end == p) throw duckdb::InvalidInputException("value json: expected a number");
  // integers stay INTEGER values, everything else is a FLOAT value (the STRUCT's declared leaf types)
  bool integral = true;
  for (const char *q = p; q < end; q++) integral = integral && (*q == '-' || (*q >= '0' && *q <= '9'));
  p = end;
  return integral ? duckdb::Value((int32_t)d) : duckdb::Value((float)d);
}

void Render(const duckdb::Value &v, std::string &out) {
  if (v.IsStruct()) {
    out += '{';
    for (size_t i = 0; i < v.NestedChildren().size(); i++) {
      if (i) out += ',';
      out += '"' + v.ChildNames()[i] + "\":";
      Render(v.NestedChildren()[i], out);
    }
    out += '}';
  } else if (v.IsList()) {
    out += '[';
    for (size_t i = 0; i < v.NestedChildren().size(); i++) {
      if (i) out += ',';
      Render(v.NestedChildren()[i], out);
    }
    out += ']';
  } else {
    char buf[40];
    snprintf(buf, sizeof buf, "%.17g", v.GetValue<double>());
    out += buf;
  }
}
}  // namespace

extern "C" {
/* op: 0 = sum_triple(a, b), 1 = subtract_triple(a, b), 2 = sum_nb_triple(a, b).  *json_out is malloc'd (free with
 * replay_free).  Returns 0, or -1 with the exception text in replay_value_error().                               */
int replay_value_ring(int op, const char *json_a, const char *json_b, char **json_out) {
  try {
    duckdb::Value a = Parse(json_a), b = Parse(json_b);
    duckdb::Value r = op == 0 ? Triple::sum_triple(a, b) : op == 1 ? Triple::subtract_triple(a, b) : Triple::sum_nb_triple(a, b);
    std::string s;
    Render(r, s);
    *json_out = strdup(s.c_str());
    return 0;
  } catch (const std::exception &e) {
    g_value_error = e.what();
    return -1;
  }
}
const char *replay_value_error(void) { return g_value_error.c_str(); }
}
