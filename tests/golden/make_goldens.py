"""Extract the golden STRUCTs of the reference's own tests into reference_goldens.json.

Run in the build container (needs /root/reference); the JSON is committed and is what the
test-suite reads -- /root/reference does not exist on the GPU box.

    python tests/golden/make_goldens.py

For every `duckdb_conn.execute("<SQL>")` followed by `assert(res[i][0] == eval("<dict>"))` in
duckdb_extension/test/python/{test_sum,test_nb_sum,test_lift,test_nb_lift,test_mul,test_nb_mul}.py the script
records {file, test, sql, index, expected}.  The fixture tables (CREATE TABLE / INSERT) are
recorded too.
"""
import ast
import json
import os
import re

REF = "/root/reference/duckdb_extension/test/python"
FILES = ["test_sum.py", "test_nb_sum.py", "test_lift.py", "test_nb_lift.py", "test_mul.py", "test_nb_mul.py"]


def main():
    out = {"source": "eddbase/duckdb-imputation duckdb_extension/test/python", "fixtures": {}, "cases": []}
    for fn in FILES:
        src = open(os.path.join(REF, fn)).read()
        ins = re.search(r'INSERT INTO test VALUES (.*?)"\)', src).group(1)
        rows = ast.literal_eval("[" + ins + "]")
        cols = re.search(r"CREATE TABLE test\((.*?)\);", src).group(1)
        out["fixtures"][fn] = {"columns": [c.strip().split()[0] for c in cols.split(",")],
                               "types": [c.strip().split()[1] for c in cols.split(",")],
                               "rows": [list(r) for r in rows]}
        for m in re.finditer(r"def (test_\w+)\s*\(duckdb_conn\):(.*?)(?=\ndef |\Z)", src, re.S):
            name, body = m.group(1), m.group(2)
            sql = None
            body = re.sub(r'"\s*\n\s*"', "", body)  # adjacent string literals (test_mul.py splits its SQL)
            for line in body.splitlines():
                e = re.search(r'execute\("(.*)"\)', line)
                if e:
                    sql = e.group(1)
                a = re.search(r'assert\s*\(res\[(\d+)\]\[0\] == eval\("(.*)"\)\)', line)
                if a:
                    out["cases"].append({"file": fn, "test": name, "sql": sql, "index": int(a.group(1)),
                                         "expected": ast.literal_eval(a.group(2))})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
    json.dump(out, open(path, "w"), indent=1)
    print(f"{len(out['cases'])} golden cases -> {path}")


if __name__ == "__main__":
    main()
