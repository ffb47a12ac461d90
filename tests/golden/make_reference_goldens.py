"""Writes tests/golden/reference_goldens.json.

The reference is Julia and cannot run in the build image, so these goldens are
TRANSCRIBED (not generated) from the reference's own tests and jldoctests; each
entry cites where the value is recorded (paths relative to the reference
checkout).  Re-run only to re-serialise.
"""
import json, os

G = {
  "_source": "JuliaPhylo/PhyloGaussianBeliefProp.jl v0.0.1 test/ and docs/ known answers",
  # test/test_calibration.jl:3-4 ; test/test_canonicalform.jl:3 ; test/test_clustergraph.jl:2-4
  "netstr_unnamed": "(A:2.5,((B:1,#H1:0.5::0.1):1,(C:1,(D:0.5)#H1:0.5::0.9):1):0.5);",
  "netstr_named": "(((A:4.0,((B1:1.0,B2:1.0)i6:0.6)#H5:1.1::0.9)i4:0.5,(#H5:2.0::0.1,C:0.1)i2:1.0)i1:3.0);",
  "netstr_cg": "(((A:4.0,(B:1.0)#H1:1.1::0.9):0.5,((#H1:1.0::0.1,C:0.6):1.0,C2):1.0):3.0,D:5.0);",
  "mateescu": "((((g:1)#H4:1)#H2:2.04,(d:1,(#H2:0.01::0.5,#H4:1::0.5)#H3:1)D:1,(#H3:1::0.5)#H1:0.01)B:1,#H1:1.01::0.5)A;",
  # test/test_calibration.jl:132
  "netstr_level3": "((#H1:0.1::0.4,#H2:0.1::0.4)I1:1.0,(((A:1.0)#H1:0.1::0.6,#H3:0.1::0.4)#H2:0.1::0.6,(B:1.0)#H3:0.1::0.6)I2:1.0)I3;",
  # test/example_networks/lazaridis_2014.phy (also printed at docs/src/man/getting_started.md:39)
  "lazaridis": "(Mbuti:1.0,(((Onge:1.0,#H1:0.01::0.4)EasternNorthAfrican:1.0,(((Karitiana:1.0)#H1:0.01::0.6,(MA1:1.0,#H3:0.01::0.4)ANE:1.0)AncientNorthEurasian:1.0,(((#H2:0.01::0.4)#H3:0.01::0.6,Loschbour:1.0)WHG:1.0,#H4:0.01::0.4)WestEurasian:1.0)I1:1.0)I2:1.0,((European:1.0)#H2:0.01::0.6,Stuttgart:1.0)#H4:0.01::0.6)NonAfrican:1.0)I3;",
  # test/test_evomodels.jl:156
  "preorder_named": ["i1","i2","C","i4","H5","i6","B2","B1","A"],
  # test/test_clustergraph.jl:11-12, 122-123, 102-106
  "minfill_order_cg": ["A","B","H1","C","C2","D","I5","I1","I2","I3","I4"],
  "cliquetree_sepsets_cg": [[1],[3],[4],[6],[6,3],[8],[8,6],[9]],
  "jgs3_clusters_mateescu": [[1],[2,1],[3,2,1],[4,3,2],[5,2],[5,4,3],[6,5,2],[7,6,5],[8,7],[9,4]],
  "jgs3_sepsets_mateescu": [[1],[2],[2,1],[3,2],[4],[4,3],[5],[5,2],[6,5],[7]],
  # test/test_canonicalform.jl:55, 109
  "canonicalform_beliefnodelabels": [[6,5],[7,6],[8,6],[5,4,2],[4,2,1],[3,2],[9,4],[6],[6],[5],[4,2],[2],[4]],
  "canonicalform_loglik": -10.732857817537196,
  # docs/src/man/getting_started.md:107-125, 46-58, 184-189, 245-261, 283-291
  "lazaridis_cluster_labels": ["H1EasternNorthAfricanAncientNorthEurasian","EasternNorthAfricanAncientNorthEurasianI2","OngeEasternNorthAfrican","StuttgartH4","MbutiI3","H2H3H4","H3ANEWHGH4","ANEWHGH4WestEurasian","LoschbourWHG","KaritianaH1","EuropeanH2","AncientNorthEurasianWestEurasianI1NonAfrican","ANEH4WestEurasianNonAfrican","ANEAncientNorthEurasianWestEurasianNonAfrican","AncientNorthEurasianI1I2NonAfrican","NonAfricanI3","MA1ANE"],
  "lazaridis_x": [1.343, 0.841, -0.623, -1.483, 0.456, -0.081, 1.311],
  "lazaridis_b1_J": [[192.30769230769232,-76.92307692307693,-115.38461538461539],[-76.92307692307693,30.769230769230774,46.15384615384616],[-115.38461538461539,46.15384615384616,69.23076923076923]],
  "lazaridis_b1_g": 1.7106097934927051,
  "lazaridis_sched_parent": [16,12,14,13,8,7,7,16,7,6,6,12,15,2,1,1],
  "lazaridis_sched_child": [12,14,13,8,7,9,17,5,6,11,4,15,2,1,3,10],
  "lazaridis_norm": -11.273958980921247,
  "lazaridis_fe": -11.273958980921261,
  # test/example_networks/lipson_2020b.phy ; docs/src/man/regularization.md:150-201 (trait x in tiplabels order, the two
  # ill-defined messages of an unregularised iteration on the Bethe graph, none after either regularisation)
  "lipson": "((((#H1:0.01::0.04,Altai:0.71)I2:0.16,((South_Africa_HG:0.16,((#H2:0.01::0.28,#H4:0.01::0.25)I3:0.01,((((French:0.12,((Agaw:0.01)#H3:0.01::0.9)#H2:0.01::0.72)I1:0.21)#H1:0.01::0.96,((#H3:0.01::0.1,Mota:1.3)I4:0.15)#H5:0.01::0.58)I5:0.16,((Cameroon_SMA:0.08)#H7:0.01::0.68,((((Biaka:0.03)#H8:0.01::0.62,#H9:0.01::0.3)I6:1.0,Lemande:0.01)I7:0.01,(Yoruba:0.01,(Mende:0.01)#H10:0.01::0.97)I8:0.01)I9:0.03)#H6:0.01::0.84)I10:0.04)I11:0.28)I12:0.04,(((#H8:0.01::0.38,#H7:0.01::0.32)I13:0.07,((Mbuti:0.1)#H4:0.01::0.75)#H9:0.01::0.7)I14:0.13,(#H5:0.01::0.42,((#H6:0.01::0.16,#H10:0.01::0.03)I15:0.01)#H11:0.01::0.7)I16:0.53)I17:0.01)I18:0.5)I19:0.09,#H11:0.01::0.3)I20:0.5,Chimp:0.5)I21;",
  "lipson_taxa": ["Altai","South_Africa_HG","French","Agaw","Mota","Cameroon_SMA","Biaka","Lemande","Yoruba","Mende","Mbuti","Chimp"],
  "lipson_x": [0.431,1.606,0.72,0.944,0.647,1.263,0.46,1.079,0.877,0.748,1.529,-0.469],
  "lipson_bethe_failing_beliefs": [["H5I5I16", [2, 3]], ["H10I8I15", [2, 3]]],
  # test/test_evomodels.jl:85,96,107,120,180,190,200,212,223,234,247,262
  "evomodels": [
    {"id":"uniBM_fixed_y","model":"UnivariateBrownianMotion","args":"(2,3,0)","traits":"y","loglik":-10.732857817537196},
    {"id":"uniBM_improper_y","model":"UnivariateBrownianMotion","args":"(2,3,inf)","traits":"y","loglik":-5.899094849099194},
    {"id":"uniBM_random_x_missing","model":"UnivariateBrownianMotion","args":"(2,3,0.4)","traits":"x","loglik":-13.75408386332493},
    {"id":"uniOU_random_y","model":"UnivariateOrnsteinUhlenbeck","args":"(2,3,-2,0.0,0.4)","traits":"y","loglik":-42.31401134496844},
    {"id":"diagBM_fixed","model":"MvDiagBrownianMotion","args":"([2,1],[3,-3],[0,0])","traits":"xy","loglik":-24.8958130127972},
    {"id":"diagBM_random","model":"MvDiagBrownianMotion","args":"([2,1],[3,-3],[0.1,10])","traits":"xy","loglik":-21.347496753649892},
    {"id":"diagBM_improper","model":"MvDiagBrownianMotion","args":"([2,1],[1,-3],[inf,inf])","traits":"xy","loglik":-17.66791635814575},
    {"id":"fullBM_fixed","model":"MvFullBrownianMotion","args":"([[2.0,0.5],[0.5,1.0]],[3.0,-3.0])","traits":"xy","loglik":-24.312323855394055},
    {"id":"fullBM_random","model":"MvFullBrownianMotion","args":"([[2.0,0.5],[0.5,1.0]],[3.0,-3.0],[[0.1,0.01],[0.01,0.2]])","traits":"xy","loglik":-23.16482738327936},
    {"id":"fullBM_improper","model":"MvFullBrownianMotion","args":"([[2.0,0.5],[0.5,1.0]],[3.0,-3.0],[[inf,0],[0,inf]])","traits":"xy","loglik":-16.9626044836951},
    {"id":"heteroBM_fixed_onerate","model":"HeterogeneousBrownianMotion","args":"([[[2.0,0.5],[0.5,1.0]]],None,[3.0,-3.0])","traits":"xy","loglik":-24.312323855394055},
    {"id":"heteroBM_random_tworates","model":"HeterogeneousBrownianMotion","args":"([[[2.0,0.5],[0.5,1.0]],[[2.0,0.5],[0.5,1.0]]],{9:2,7:2,8:2},[3.0,-3.0],[[0.1,0.01],[0.01,0.2]])","traits":"xy","loglik":-23.16482738327936}
  ]
}
json.dump(G, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json"), "w"), indent=1)
