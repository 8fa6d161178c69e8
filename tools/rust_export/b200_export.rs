//! Dumps the SHA-256 bench circuit of this crate (setup, witness, gate programs, VK, reference proof) for the B200 backend.
//! Drop into src/gadgets/sha256/ and add `#[cfg(test)] mod b200_export;` to src/gadgets/sha256/mod.rs.  See README.md next to
//! this file in the boojum-b200 repository for the file formats.  Only calls items of this crate:
//!   circuit + configuration     src/gadgets/sha256/mod.rs:296-470 (same geometry, lookup tables and gate order)
//!   get_full_setup              src/cs/implementations/setup.rs:1273-1298
//!   take_witness_using_hints    src/cs/implementations/prover.rs:119-151
//!   prove_cpu_basic             src/cs/implementations/prover.rs:153-168
//!   GPUDataCapture              src/gpu_synthesizer/mod.rs:354-443
#![allow(clippy::all)]
use std::alloc::Global;
use std::io::Write;

use crate::config::*;
use crate::cs::cs_builder::*;
use crate::cs::cs_builder_reference::*;
use crate::cs::gates::*;
use crate::cs::implementations::pow::NoPow;
use crate::cs::implementations::prover::ProofConfig;
use crate::cs::implementations::transcript::*;
use crate::cs::oracle::TreeHasher;
use crate::cs::traits::cs::ConstraintSystem;
use crate::cs::traits::evaluator::*;
use crate::cs::traits::gate::*;
use crate::cs::*;
use crate::dag::CircuitResolverOpts;
use crate::field::goldilocks::{GoldilocksExt2, GoldilocksField};
use crate::field::traits::field_like::PrimeFieldLikeVectorized;
use crate::field::{Field, PrimeField, U64Representable};
use crate::gadgets::sha256::sha256;
use crate::gadgets::tables::*;
use crate::gadgets::traits::allocatable::CSAllocatable;
use crate::gadgets::u8::UInt8;
use crate::gpu_synthesizer::{GPUDataCapture, Index, Relation};
use crate::worker::Worker;

type F = GoldilocksField;
type P = GoldilocksField; // scalar field-like: the storages are then plain Vec<GoldilocksField>

fn out_dir() -> std::path::PathBuf {
    let d = std::env::var("B200_EXPORT_DIR").unwrap_or_else(|_| "b200_dump".to_string());
    std::fs::create_dir_all(&d).unwrap();
    std::path::PathBuf::from(d)
}

fn write_columns<'a, I: Iterator<Item = &'a [F]>>(name: &str, cols: I) -> usize {
    let mut f = std::io::BufWriter::new(std::fs::File::create(out_dir().join(name)).unwrap());
    let mut n = 0;
    for col in cols {
        for el in col.iter() {
            f.write_all(&el.as_u64_reduced().to_le_bytes()).unwrap();
        }
        n += 1;
    }
    n
}

fn index_json(idx: &Index<F>, shared: &[Index<F>]) -> String {
    match idx {
        Index::VariablePoly(i) => format!("[\"variable\",{}]", i),
        Index::WitnessPoly(i) => format!("[\"witness\",{}]", i),
        Index::ConstantPoly(i) => {
            // a constant that load_row_shared_constants read is the same for every repetition on the row
            let kind = if shared.contains(idx) { "constant_shared" } else { "constant" };
            format!("[\"{}\",{}]", kind, i)
        }
        Index::TemporaryValue(i) => format!("[\"temporary\",{}]", i),
        Index::ConstantValue(v) => format!("[\"value\",{}]", v.as_u64_reduced()),
    }
}

fn capture_json<E: GateConstraintEvaluator<F>>(gate_idx: usize, evaluator: E, geometry: &CSGeometry) -> String {
    let reps = evaluator.num_repetitions_in_geometry(geometry);
    let offsets = evaluator.per_chunk_offset_for_repetition_over_general_purpose_columns();
    let terms = E::num_quotient_terms();
    let cap = GPUDataCapture::from_evaluator(evaluator);
    let shared: Vec<Index<F>> = cap.row_shared_constants_set.iter().map(|el| el.idx).collect();
    let mut rel = vec![];
    for (dst, r) in cap.relations.iter() {
        let (op, a, b) = match r {
            Relation::Add(a, b) => ("add", a, Some(b)),
            Relation::Sub(a, b) => ("sub", a, Some(b)),
            Relation::Mul(a, b) => ("mul", a, Some(b)),
            Relation::Double(a) => ("double", a, None),
            Relation::Negate(a) => ("negate", a, None),
            Relation::Square(a) => ("square", a, None),
            Relation::Inverse(a) => ("inverse", a, None),
        };
        let b = b.map(|b| index_json(b, &shared)).unwrap_or_else(|| "null".to_string());
        rel.push(format!("[\"{}\",{},{},{}]", op, index_json(&dst.idx, &shared), index_json(a, &shared), b));
    }
    let writes: Vec<String> = cap.writes_per_repetition.iter().map(|w| index_json(w, &shared)).collect();
    format!(
        "{{\"name\":\"{}\",\"gate_idx\":{},\"num_repetitions\":{},\"num_quotient_terms\":{},\"variables_offset\":{},\"witnesses_offset\":{},\"constants_offset\":{},\"relations\":[{}],\"writes\":[{}]}}",
        cap.evaluator_name, gate_idx, reps, terms, offsets.variables_offset, offsets.witnesses_offset, offsets.constants_offset,
        rel.join(","), writes.join(",")
    )
}

fn export<T: TreeHasher<F, Output = TR::CompatibleCap>, TR: Transcript<F, TransciptParameters = ()>>(hasher: &str, transcript: &str)
where
    T::Output: serde::Serialize + serde::de::DeserializeOwned,
{
    let len: usize = std::env::var("B200_EXPORT_LEN").ok().and_then(|s| s.parse().ok()).unwrap_or(8 * (1 << 10));
    let worker = Worker::new_with_num_threads(8);
    let (lde, cap_size) = (8usize, 16usize);
    let prover_config = ProofConfig { fri_lde_factor: lde, pow_bits: 0, ..Default::default() };
    use rand::{Rng, SeedableRng};
    let mut rng = rand::rngs::StdRng::seed_from_u64(42);
    let input: Vec<u8> = (0..len).map(|_| rng.gen()).collect();
    let geometry = CSGeometry {
        num_columns_under_copy_permutation: 60,
        num_witness_columns: 0,
        num_constant_columns: 4,
        max_allowed_constraint_degree: 4,
    };
    let max_variables = 1 << 27;
    let max_trace_len = std::cmp::max(1 << 19, (len * 8).next_power_of_two()); // 2^22 rows need ~512 KiB of input

    fn configure<TI: CsBuilderImpl<F, TI>, GC: GateConfigurationHolder<F>, TB: StaticToolboxHolder>(
        builder: CsBuilder<TI, F, GC, TB>,
    ) -> CsBuilder<TI, F, impl GateConfigurationHolder<F>, impl StaticToolboxHolder> {
        let builder = builder.allow_lookup(LookupParameters::UseSpecializedColumnsWithTableIdAsConstant {
            width: 4,
            num_repetitions: 8,
            share_table_id: true,
        });
        let builder = ConstantsAllocatorGate::configure_builder(builder, GatePlacementStrategy::UseGeneralPurposeColumns);
        let builder = FmaGateInBaseFieldWithoutConstant::configure_builder(builder, GatePlacementStrategy::UseGeneralPurposeColumns);
        let builder = ReductionGate::<F, 4>::configure_builder(builder, GatePlacementStrategy::UseGeneralPurposeColumns);
        NopGate::configure_builder(builder, GatePlacementStrategy::UseGeneralPurposeColumns)
    }
    macro_rules! synthesize {
        ($cs:expr) => {{
            $cs.add_lookup_table::<TriXor4Table, 4>(create_tri_xor_table());
            $cs.add_lookup_table::<Ch4Table, 4>(create_ch4_table());
            $cs.add_lookup_table::<Maj4Table, 4>(create_maj4_table());
            $cs.add_lookup_table::<Split4BitChunkTable<1>, 4>(create_4bit_chunk_split_table::<F, 1>());
            $cs.add_lookup_table::<Split4BitChunkTable<2>, 4>(create_4bit_chunk_split_table::<F, 2>());
            let bytes: Vec<_> = input.iter().map(|b| UInt8::allocate_checked(&mut $cs, *b)).collect();
            let _ = sha256(&mut $cs, &bytes);
        }};
    }

    // setup pass
    let builder = configure(new_builder::<_, F>(CsReferenceImplementationBuilder::<F, P, SetupCSConfig>::new(geometry, max_trace_len)));
    let mut cs = builder.build(CircuitResolverOpts::new(max_variables));
    synthesize!(cs);
    let (_, padding_hint) = cs.pad_and_shrink();
    let cs = cs.into_assembly::<Global>();
    let (base_setup, setup, vk, setup_tree, vars_hint, wits_hint) = cs.get_full_setup::<T>(&worker, lde, cap_size);

    // proving pass
    let builder = configure(new_builder::<_, F>(CsReferenceImplementationBuilder::<F, P, ProvingCSConfig>::new(geometry, max_trace_len)));
    let mut cs = builder.build(CircuitResolverOpts::new(max_variables));
    synthesize!(cs);
    cs.pad_and_shrink_using_hint(&padding_hint);
    let mut cs = cs.into_assembly::<Global>();
    let witness_set = cs.take_witness_using_hints(&worker, &vars_hint, &wits_hint);

    // ---- dump ----
    let n_vars = write_columns("variables.bin", witness_set.variables.iter().map(|p| &p.storage[..]));
    let n_wits = write_columns("witness.bin", witness_set.witness.iter().map(|p| &p.storage[..]));
    let n_mult = write_columns("multiplicities.bin", witness_set.multiplicities.iter().map(|p| &p.storage[..]));
    let n_sig = write_columns("sigmas.bin", base_setup.copy_permutation_polys.iter().map(|p| &p.storage[..]));
    let n_const = write_columns("constants.bin", base_setup.constant_columns.iter().map(|p| &p.storage[..]));
    let n_tab = write_columns("tables.bin", base_setup.lookup_tables_columns.iter().map(|p| &p.storage[..]));
    assert_eq!(n_vars, n_sig);
    // gate programs in registration order of the evaluators over general-purpose columns (the index the selector tree uses)
    let gates = vec![
        capture_json(0, <ConstantsAllocatorGate<F> as Gate<F>>::Evaluator::new_from_parameters(()), &geometry),
        capture_json(1, <FmaGateInBaseFieldWithoutConstant<F> as Gate<F>>::Evaluator::new_from_parameters(()), &geometry),
        capture_json(2, <ReductionGate<F, 4> as Gate<F>>::Evaluator::new_from_parameters(()), &geometry),
    ];
    let fp = &vk.fixed_parameters;
    let manifest = format!(
        "{{\"domain_size\":{},\"num_variables\":{},\"num_witness\":{},\"num_multiplicities\":{},\"num_constants\":{},\"num_tables\":{},\
         \"geometry\":{},\"lookup_parameters\":{},\"quotient_degree\":{},\"fri_lde_factor\":{},\"cap_size\":{},\"security_level\":{},\"pow_bits\":{},\
         \"hasher\":\"{}\",\"transcript\":\"{}\",\"table_ids_column_idxes\":{},\"selectors_placement\":{},\"public_inputs_locations\":{},\
         \"extra_constant_polys_for_selectors\":{},\"gates\":[{}]}}",
        fp.domain_size, n_vars, n_wits, n_mult, n_const, n_tab,
        serde_json::to_string(&fp.parameters).unwrap(), serde_json::to_string(&fp.lookup_parameters).unwrap(),
        fp.quotient_degree, lde, cap_size, prover_config.security_level, prover_config.pow_bits, hasher, transcript,
        serde_json::to_string(&fp.table_ids_column_idxes).unwrap(), serde_json::to_string(&fp.selectors_placement).unwrap(),
        serde_json::to_string(&fp.public_inputs_locations).unwrap(), fp.extra_constant_polys_for_selectors, gates.join(",")
    );
    std::fs::write(out_dir().join("manifest.json"), manifest).unwrap();
    std::fs::write(out_dir().join("vk.json"), serde_json::to_string(&vk).unwrap()).unwrap();

    // the reference proof on exactly this witness
    let proof = cs.prove_cpu_basic::<GoldilocksExt2, TR, T, NoPow>(&worker, witness_set, &base_setup, &setup, &setup_tree, &vk, prover_config, ());
    std::fs::write(out_dir().join("proof.json"), serde_json::to_string(&proof).unwrap()).unwrap();
}

#[test]
#[ignore]
fn b200_export_non_recursive() {
    use crate::blake2::Blake2s256;
    export::<Blake2s256, Blake2sTranscript>("blake2s", "blake2s");
}

#[test]
#[ignore]
fn b200_export_recursive_poseidon2() {
    use crate::algebraic_props::round_function::AbsorptionModeOverwrite;
    use crate::algebraic_props::sponge::GoldilocksPoseidon2Sponge;
    export::<GoldilocksPoseidon2Sponge<AbsorptionModeOverwrite>, GoldilocksPoisedonTranscript>("poseidon2", "poseidon");
}
