function [Floe,dissolvedNEW] = floe_interactions_all(Floe, floebound, ocean, winds,c2_boundary, dt, HFo, min_floe_size, Nx,Ny,Nb, dissolvedNEW,doInt,COLLISION, PERIODIC, RIDGING, RAFTING)
%FLOE_INTERACTIONS_ALL  Drop-in for the reference's floe_interactions_all.m (same name, same signature,
% floe_interactions_all.m:1) whose contact range runs on a B200 through libsubzero_b200.so.
%
%   reference lines   what happens here
%   :3-14             warnings off, Lx/Ly, dead floes dropped                        -- below, in MATLAB
%   :16-277           ghost floes, potential interactions, floe_interactions over every pair and the walls,
%                     kill/transfer, mirror, torques, force/torque sums, periodic wrap -- ONE call: sz_contact_step -> sz_contact_mex
%   :279-283          the reference's own calc_trajectory per floe and its kill rule   -- below, unchanged calls
%   :16-66 (structs)  ghost structs Floe(N0+1:N): only the ridging / rafting tail reads them (Floe(partner), partner > N0,
%                     :312,327,401,416), so they are materialised only when that tail will run
%   :286-512          ridging, rafting, kill / fuse, rmfield                          -- the reference's own code, executed from
%                     sz_floe_interactions_tail.m, which sz_make_tail.m cuts out of YOUR copy of the reference file
%                     (nothing of the reference is shipped with this repository)
%
% Install (INTEGRATION.md section 2): rename the reference's floe_interactions_all.m to floe_interactions_all_reference.m,
% run sz_make_tail('<SubZero>/floe_interactions_all_reference.m') once, and put this folder on the path ahead of SubZero's.

id ='MATLAB:polyshape:repairedBySimplify';          % the tail switches both back on (:509-510)
warning('off',id)
id3 ='MATLAB:polyshape:boundary3Points';
warning('off',id3)
global Modulus

c2_boundary_poly = polyshape(c2_boundary');         % read by the tail (ridge / raft / calc_dissolved_mass)
live = cat(1,Floe.alive);
Floe(live==0)=[];                                   % :12-13
N0 = length(Floe);

% pre-step centroids: the kill rule of :282 tests isnan(x(i)) on the centroid gathered BEFORE the wrap (:71)
x = cat(1,Floe.Xi);
alive0 = cat(1,Floe.alive);

% ---- :16-277 on the GPU.  On return every floe i > Nb carries interactions (K x 7, torque in column 6), OverlapArea,
% collision_force, collision_torque, Stress = zeros(2), potentialInteractions = [], alive, wrapped Xi / Yi.
[Floe, kill, transfer, ghosts] = sz_contact_step(Floe, floebound, c2_boundary, dt, Nb, COLLISION, PERIODIC, Modulus);

% ---- ghost structs for the tail (:16-66 as data).  A ghost is its parent as it was when the step began, with the shifted
% centroid, and carries what :76-88,218-238 leave in it.  Built before the integrator moves the parents.
need_ghosts = PERIODIC && doInt.flag && (RIDGING || RAFTING) && ~isempty(ghosts.parent);
if need_ghosts
    G = repmat(Floe(1), 1, numel(ghosts.parent));
    for k = 1:numel(ghosts.parent)
        p = ghosts.parent(k);
        if p <= N0
            g = Floe(p);  g.alive = alive0(p);               % the parent as it was before the wall test
        else
            g = G(p - N0);                                   % a y-ghost of an x-ghost (:49-60 runs over the extended list)
        end
        g.Xi = ghosts.x(k);  g.Yi = ghosts.y(k);             % shifted centroid (:34,55)
        g.interactions = ghosts.interactions{k};
        g.OverlapArea = ghosts.overlap_area(k);
        g.collision_force = [ghosts.fx(k) ghosts.fy(k)];
        g.collision_torque = ghosts.torque(k);
        g.Stress = zeros(2);
        g.potentialInteractions = [];
        G(k) = g;
    end
end

% ---- :279-283, the reference's integrator and kill rule (calc_trajectory stays the reference's function)
for i=1+Nb:N0
    if Floe(i).alive
        [tmp,Fx,Fy] = calc_trajectory(dt,ocean,winds,Floe(i),HFo,doInt); %#ok<ASGLU>
        if (isempty(tmp) || isnan(x(i)) ), kill(i)=i; else; Floe(i)=tmp; end
    end
end

if need_ghosts
    Floe = [Floe G];                                         % the tail drops them again at :468
end

% ---- :286-512, the reference's own tail
[Floe,dissolvedNEW] = sz_floe_interactions_tail(Floe, floebound, c2_boundary, c2_boundary_poly, min_floe_size, Nx, Ny, Nb, N0, ...
                                                kill, transfer, dissolvedNEW, doInt, PERIODIC, RIDGING, RAFTING, id, id3);
end
