//! Golden-vector dump for the B200 prover's parity kit (see README.md next to this file).
//! Drop into `recursion/examples/common/` of the reference and call `dump(label, &proof)` on a `BatchStarkProof`.
//! SOURCE ONLY: never compiled in the repository that ships it (no Rust toolchain there).

use p3_field::extension::BinomialExtensionField;
use p3_field::{BasedVectorSpace, PrimeCharacteristicRing, PrimeField32};
use serde::Serialize;
use serde_json::json;

fn hex(bytes: &[u8]) -> String {
    const DIGITS: &[u8; 16] = b"0123456789abcdef";
    let mut s = String::with_capacity(bytes.len() * 2);
    for b in bytes {
        s.push(DIGITS[(b >> 4) as usize] as char);
        s.push(DIGITS[(b & 15) as usize] as char);
    }
    s
}

/// How the base field and its degree-4 extension serialise: tells Montgomery from canonical form.
pub fn probe<F>() -> serde_json::Value
where
    F: PrimeField32 + Serialize,
    BinomialExtensionField<F, 4>: Serialize + BasedVectorSpace<F>,
{
    let two = F::TWO;
    let minus_one = -F::ONE;
    let ext = BinomialExtensionField::<F, 4>::from_basis_coefficients_fn(|i| F::from_usize(i + 1));
    json!({
        "order": F::ORDER_U32,
        "two_serde": serde_json::to_value(two).unwrap(),
        "two_postcard_hex": hex(&postcard::to_allocvec(&two).unwrap()),
        "minus_one_serde": serde_json::to_value(minus_one).unwrap(),
        "minus_one_canonical": minus_one.as_canonical_u32(),
        "ext_1_2_3_4_serde": serde_json::to_value(ext).unwrap(),
        "ext_1_2_3_4_postcard_hex": hex(&postcard::to_allocvec(&ext).unwrap()),
    })
}

/// Write `ref_<label>.json`: the proof as postcard bytes (what `report_proof_size` measures) and, field by field and by name,
/// through serde_json.
pub fn dump<S: Serialize>(label: &str, proof: &S) {
    let bytes = postcard::to_allocvec(proof).expect("postcard");
    let value = json!({
        "label": label,
        "postcard_len": bytes.len(),
        "postcard_hex": hex(&bytes),
        "serde": serde_json::to_value(proof).expect("serde_json"),
        "probe_koala_bear": probe::<p3_koala_bear::KoalaBear>(),
        "probe_baby_bear": probe::<p3_baby_bear::BabyBear>(),
    });
    let path = format!("ref_{label}.json");
    std::fs::write(&path, serde_json::to_vec_pretty(&value).expect("json")).expect("write");
    println!("ref_golden: wrote {path} ({} proof bytes)", bytes.len());
}
