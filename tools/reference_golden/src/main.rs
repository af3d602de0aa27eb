//! Runs stock CozoDB (mem engine) on the inputs the oracle tests use and dumps what it returns:
//!   reference_hnsw.json   config 1 (10k x 128 f32, m=16, ef_construction=200): the index relation rows
//!                         (so the oracle / the device can be staged from the REFERENCE's own graph — its
//!                         levels come from thread_rng and cannot be seeded) and, for every query, the rows
//!                         `~a:vec{ k | query: q, k: 10, ef: 64, bind_distance: d }` returns;
//!   reference_graph.json  air-routes: PageRank, ShortestPathDijkstra (JFK/LHR/SYD/AUS/KUL to all),
//!                         ClosenessCentrality, BetweennessCentrality.
//! Inputs: argv[1] = directory with hnsw_vectors.json / hnsw_queries.json (written by
//! tests/golden/make_reference_inputs.py), argv[2] = cozo-core/tests (the air-routes CSVs), argv[3] = out dir.
use cozo::{DataValue, DbInstance, ScriptMutability};
use serde_json::{json, Value};
use std::collections::BTreeMap;

fn rows_json(r: &cozo::NamedRows) -> Value {
    json!({"headers": r.headers, "rows": r.rows.iter().map(|t| t.iter().map(|d| serde_json::to_value(d).unwrap()).collect::<Vec<_>>()).collect::<Vec<_>>()})
}

fn main() {
    let a: Vec<String> = std::env::args().collect();
    let (inp, csv, out) = (&a[1], &a[2], &a[3]);
    let db = DbInstance::new("mem", "", "").unwrap();
    // ---- HNSW, config 1 ------------------------------------------------------------------------------
    let vectors: Value = serde_json::from_str(&std::fs::read_to_string(format!("{inp}/hnsw_vectors.json")).unwrap()).unwrap();
    let queries: Value = serde_json::from_str(&std::fs::read_to_string(format!("{inp}/hnsw_queries.json")).unwrap()).unwrap();
    db.run_default(":create a {k: Int => v: <F32; 128>}").unwrap();
    let mut p = BTreeMap::new();
    p.insert("rows".to_string(), DataValue::from(vectors));
    db.run_script("?[k, v] <- $rows :put a {k => v}", p, ScriptMutability::Mutable).unwrap();
    db.run_default("::hnsw create a:vec {dim: 128, m: 16, dtype: F32, fields: [v], distance: L2, ef_construction: 200, extend_candidates: false, keep_pruned_connections: false}").unwrap();
    let idx = db.run_default("?[layer, fr_k, fr__field, fr__sub_idx, to_k, to__field, to__sub_idx, dist, hash, ignore_link] := *a:vec{layer, fr_k, fr__field, fr__sub_idx, to_k, to__field, to__sub_idx, dist, hash, ignore_link}").unwrap();
    let mut results = vec![];
    for q in queries.as_array().unwrap() {
        let mut p = BTreeMap::new();
        p.insert("q".to_string(), DataValue::from(q.clone()));
        let r = db.run_script("?[k, d] := ~a:vec{k | query: q, k: 10, ef: 64, bind_distance: d}, q = vec($q)", p, ScriptMutability::Immutable).unwrap();
        results.push(rows_json(&r));
    }
    std::fs::write(format!("{out}/reference_hnsw.json"), serde_json::to_string(&json!({"index_rows": rows_json(&idx), "knn": results})).unwrap()).unwrap();
    // ---- graph rules on air-routes (cozo-core/tests/air_routes.rs:96-127) ---------------------------------
    db.run_default(&format!("res[idx, label, typ, code] <~ CsvReader(types: ['Int', 'Any', 'Any', 'Any'], url: 'file://{csv}/air-routes-latest-nodes.csv', has_headers: true) ?[idx, code] := res[idx, label, typ, code] :replace idx2code {{ idx: Int => code: String }}")).unwrap();
    db.run_default(&format!("res[] <~ CsvReader(types: ['Int', 'Int', 'Int', 'String', 'Float?'], url: 'file://{csv}/air-routes-latest-edges.csv', has_headers: true) ?[fr, to, dist] := res[idx, fr_i, to_i, typ, dist], typ == 'route', *idx2code[fr_i, fr], *idx2code[to_i, to] :replace route {{ fr: String, to: String => dist: Float }}")).unwrap();
    let mut g = serde_json::Map::new();
    for (name, script) in [
        ("pagerank", "?[code, rank] <~ PageRank(*route[fr, to])"),
        ("pagerank_1_iteration", "?[code, rank] <~ PageRank(*route[fr, to], iterations: 1)"),
        ("dijkstra", "starting[] <- [['JFK'], ['LHR'], ['SYD'], ['AUS'], ['KUL']] ?[fr, to, cost, path] <~ ShortestPathDijkstra(*route[], starting[])"),
        ("closeness", "?[code, c] <~ ClosenessCentrality(*route[fr, to, dist])"),
        ("betweenness", "?[code, c] <~ BetweennessCentrality(*route[fr, to, dist])"),
        ("clustering", "?[code, cc, tri, deg] <~ ClusteringCoefficients(*route[fr, to])"),
        ("yen", "starting[] <- [['JFK']] goal[] <- [['KUL']] ?[fr, to, cost, path] <~ KShortestPathYen(*route[], starting[], goal[], k: 5)"),
    ] {
        g.insert(name.to_string(), rows_json(&db.run_default(script).unwrap()));
    }
    std::fs::write(format!("{out}/reference_graph.json"), serde_json::to_string(&Value::Object(g)).unwrap()).unwrap();
}
